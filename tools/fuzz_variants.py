#!/usr/bin/env python
"""Random COMBINATIONS of the single-directive variants of tests/test_variants_gpu.py (two or three edits of
scenes/smoke.pbrt at once), each rendered on the device and by the reference (ref_harness render, the same per-(pixel,
sample) streams or the stock GlobalSamplers) and compared with the tests' bars.  A search for interactions the
one-at-a-time sweep cannot see; failures are printed with the edits that produced them.
python tools/fuzz_variants.py [n] [seed]"""
import json
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402
import test_variants_gpu as tv  # noqa: E402
from test_render_gpu import film_agreement  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    iile = load_product()
    iile.set_device(0)
    work = Path("/tmp/fuzz_variants")
    work.mkdir(exist_ok=True)
    names = sorted(tv.CASES)
    done = fails = skipped = 0
    while done < n:
        picks = rng.sample(names, rng.choice((2, 3)))
        text, ok = tv.SMOKE, True
        for name in picks:
            spec = tv.CASES[name]
            for old, new in [(spec[0], spec[1])] + list(spec[2:]):
                if old not in text:
                    ok = False
                    break
                text = text.replace(old, new, 1)
            if not ok:
                break
        if not ok:  # two picks edit the same directive
            continue
        src = work / "fuzz.pbrt"
        src.write_text(text)
        try:
            scene = iile.Scene.from_pbrt(src)
            flat = work / "fuzz.flat"
            scene.save_flat(flat)
            film, stats = scene.render()
        except iile.IileError as e:  # a combination the front-end refuses (e.g. dimensions of the sample arrays)
            skipped += 1
            done += 1
            print("REFUSED", picks, str(e)[:100])
            continue
        rgb = scene.resolve(film)
        stock = scene.info().setup.samplerType in (iile.SAMPLER_HALTON, iile.SAMPLER_SOBOL)
        out = work / "fuzz.pfm"
        oracle.ref_render(flat, out, sampler="stock" if stock else "pixelstream")
        ref = oracle.read_pfm(out)
        frac, relmse = film_agreement(rgb, ref) if ref.shape == rgb.shape else (0.0, 1.0)
        good = ref.shape == rgb.shape and frac >= 0.995 and relmse <= 1e-5
        done += 1
        if not good:
            fails += 1
            keep = work / f"fail_{fails}.pbrt"
            keep.write_text(text)
        print("ok  " if good else "FAIL", picks, f"{frac:.4f} {relmse:.3g}")
    print(json.dumps({"combinations": done, "failed": fails, "refused": skipped}))


if __name__ == "__main__":
    main()
