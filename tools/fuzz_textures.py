#!/usr/bin/env python
"""The textured scenes of tests/test_textures_gpu.py under random combinations of integrator (path / directlighting /
whitted), GlobalSampler (halton / sobol), camera (perspective with and without a lens, orthographic, environment), crop
window and samples per pixel, each rendered on the device and by the STOCK reference binary and compared with the
tests' bars (>= 99.5 % of the pixels within 1e-5, 99.99 % within 1e-3, relMSE <= 1e-9).  python tools/fuzz_textures.py [n] [seed]"""
import json
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402
import test_textures_gpu as tt  # noqa: E402

CAMERAS = ['Camera "perspective" "float fov" [42] ',
           'Camera "perspective" "float fov" [42] "float lensradius" [0.03] "float focaldistance" [3]',
           'Camera "orthographic" "float screenwindow" [-2.2 2.2 -1.6 1.6]',
           'Camera "environment"']


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    iile = load_product()
    iile.set_device(0)
    work = Path("/tmp/fuzz_textures")
    work.mkdir(exist_ok=True)
    fails = refused = 0
    for k in range(n):
        case = rng.choice(sorted(tt.CASES))
        src = tt._write_case(work, case)
        text = src.read_text()
        integ = rng.choice(["path", "directlighting", "whitted"])
        sampler = rng.choice(["halton", "sobol"])
        spp = rng.choice([1, 2, 4, 8])  # sobol wants a power of two
        cam = rng.choice(CAMERAS)
        crop = rng.choice(["", ' "float cropwindow" [.1 .85 .2 .9]'])
        head = text[:text.index("WorldBegin")]
        new = head
        new = new[:new.index("Camera ")] + cam + "\n" + new[new.index("Sampler "):]
        new = new[:new.index("Sampler ")] + f'Sampler "{sampler}" "integer pixelsamples" [{spp}]\n' + new[new.index("Integrator "):]
        new = new[:new.index("Integrator ")] + f'Integrator "{integ}" "integer maxdepth" [4]\n' + new[new.index("Film "):]
        new = new.replace('"string filename" "tex.pfm"', '"string filename" "tex.pfm"' + crop)
        text = new + text[text.index("WorldBegin"):]
        src.write_text(text)
        label = (case, integ, sampler, spp, cam.split('"')[1] + (" lens" if "lensradius" in cam else ""), bool(crop))
        try:
            scene = iile.Scene.from_pbrt(src)
            rgb = scene.resolve(scene.render()[0])
        except iile.IileError as e:
            refused += 1
            print("REFUSED", label, str(e)[:120])
            continue
        out = work / "ref.pfm"
        oracle._run([oracle.REF_PBRT, "--quiet", "--outfile", out, src])
        ref = oracle.read_pfm(out)
        if ref.shape != rgb.shape:
            good, f5, f3, relmse = False, 0, 0, 1
        else:
            f5, f3, relmse = tt.agreement(rgb, ref)
            good = f5 >= 0.995 and f3 >= 0.9999 and relmse <= 1e-9  # tests/test_textures_gpu.py's bar
        if not good:
            fails += 1
            (work / f"fail_{fails}.pbrt").write_text(text)
        print("ok  " if good else "FAIL", label, f"{f5:.4f} {f3:.4f} {relmse:.3g}")
    print(json.dumps({"combinations": n, "failed": fails, "refused": refused}))


if __name__ == "__main__":
    main()
