#!/usr/bin/env python
"""Whole renders of random scenes — tools/fuzz_bvh.py's triangle soups cut into chunks with random in-scope materials,
random lights (point, spot, distant, diffuse area lights on soup triangles and on quadrics, a constant infinite
light), random integrator / depth / light-sampling strategy — on the device and by the reference (ref_harness render on
the flattened scene, the same per-(pixel, sample) streams), compared with tests/test_render_gpu.py's bars.

Soups with DUPLICATED triangles in different material groups used to need a weaker bar: every ray that reaches such a
pair is an exact tie, which the traversal resolves exactly as the reference does (tools/fuzz_cast.py), but a bounce ray
whose direction carried the last ulp of CUDA's sinf / cosf instead of glibc's landed on the other twin — on another
material — in a few cases (one film at 83 %).  With the host libm's bits on the device (csrc/cuda/libm_exact.cuh) every
ray is the reference's: 120 of 120 scenes have EVERY pixel within 1e-5 (profiles/r10_fuzz_render.log).
python tools/fuzz_render.py [n] [seed]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tools"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402
from fuzz_bvh import soup  # noqa: E402
from test_render_gpu import film_agreement  # noqa: E402

MATERIALS = ['Material "matte" "color Kd" [.6 .5 .4]', 'Material "matte" "color Kd" [.3 .6 .7] "float sigma" [30]',
             'Material "plastic" "color Kd" [.4 .3 .6] "color Ks" [.3 .3 .3] "float roughness" [.15]',
             'Material "mirror"', 'Material "glass" "float index" [1.45]',
             'Material "glass" "float uroughness" [.2] "float vroughness" [.1]',
             'Material "uber" "color Kd" [.3 .3 .2] "color Ks" [.2 .2 .2] "color Kr" [.2 .2 .2] "color Kt" [.2 .2 .3]',
             'Material "substrate" "color Kd" [.4 .2 .2] "color Ks" [.1 .1 .1]', 'Material "translucent"',
             'Material "none"']


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    iile = load_product()
    iile.set_device(0)
    work = Path("/tmp/fuzz_render")
    work.mkdir(exist_ok=True)
    fails = 0
    for k in range(n):
        kind = ["plain", "scales", "lattice", "duplicates"][k % 4]
        nt = int(rng.choice([40, 200, 1500]))
        tri = soup(rng, nt, kind)
        tri[:, 1:] = tri[:, :1] + (tri[:, 1:] - tri[:, :1]) * 4      # larger triangles: something to look at
        world = ['LightSource "point" "color I" [6 6 6] "point from" [1.5 -2 2]'] if rng.random() < 0.7 else []
        if rng.random() < 0.3:
            world.append('LightSource "spot" "color I" [20 18 16] "point from" [-2 -2 2] "point to" [0 0 0] "float coneangle" [40]')
        if rng.random() < 0.3:
            world.append('LightSource "distant" "color L" [1 1 1.2] "point from" [0 -1 3] "point to" [0 0 0]')
        if rng.random() < 0.3:
            world.append('LightSource "infinite" "color L" [.3 .35 .5]')
        if rng.random() < 0.4:
            t = rng.uniform(-1, 1, 3)
            shape = rng.choice(['Shape "sphere" "float radius" [.2]', 'Shape "disk" "float radius" [.3]',
                                'Shape "cylinder" "float radius" [.1] "float zmin" [-.2] "float zmax" [.2]'])
            world.append(f'AttributeBegin\nAreaLightSource "diffuse" "color L" [9 8 7]\nTranslate {t[0]:.4f} {t[1]:.4f} {t[2] + 1.5:.4f}\n{shape}\nAttributeEnd')
        cuts = sorted(set([0, nt] + [int(c) for c in rng.integers(1, nt, size=3)]))
        for a, b in zip(cuts[:-1], cuts[1:]):
            chunk = tri[a:b]
            P = " ".join(f"{v:.9g}" for v in chunk.reshape(-1))
            idx = " ".join(str(i) for i in range(3 * len(chunk)))
            light = 'AreaLightSource "diffuse" "color L" [4 4 3] "bool twosided" "true"\n' if (rng.random() < 0.25 and len(chunk) < 60) else ""
            world.append(f'AttributeBegin\n{light}{MATERIALS[int(rng.integers(len(MATERIALS)))]}\n'
                         f'Shape "trianglemesh" "integer indices" [{idx}] "point P" [{P}]\nAttributeEnd')
        if not any("Light" in w for w in world):
            world.insert(0, 'LightSource "point" "color I" [6 6 6] "point from" [1.5 -2 2]')
        integ = rng.choice(['Integrator "path" "integer maxdepth" [%d]' % rng.integers(1, 8),
                            'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "uniform"',
                            'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "power"',
                            'Integrator "directlighting" "integer maxdepth" [4]',
                            'Integrator "directlighting" "string strategy" "one" "integer maxdepth" [3]',
                            'Integrator "whitted" "integer maxdepth" [4]'])
        text = (f'LookAt 0.3 -4 1.2  0 0 0  0 0 1\nCamera "perspective" "float fov" [38]\n'
                f'Film "image" "integer xresolution" [64] "integer yresolution" [48] "string filename" "f.pfm"\n'
                f'Sampler "random" "integer pixelsamples" [8]\n{integ}\n'
                f'Accelerator "bvh" "string splitmethod" "{rng.choice(["sah", "hlbvh", "middle", "equal"])}"\nWorldBegin\n' + "\n".join(world) + "\nWorldEnd\n")
        src = work / "s.pbrt"
        src.write_text(text)
        try:
            scene = iile.Scene.from_pbrt(src)
            flat = work / "s.flat"
            scene.save_flat(flat)
            film, stats = scene.render()
        except iile.IileError as e:
            print("REFUSED", kind, nt, str(e)[:120])
            continue
        rgb = scene.resolve(film)
        out = work / "ref.pfm"
        oracle.ref_render(flat, out, sampler="pixelstream")
        ref = oracle.read_pfm(out)
        frac, relmse = film_agreement(rgb, ref)
        good = frac >= 0.995 and relmse <= 1e-5  # one bar for every kind, duplicated triangles included
        if not good:
            fails += 1
            (work / f"fail_{fails}.pbrt").write_text(text)
        print("ok  " if good else "FAIL", kind, nt, integ.split('"')[1], f"{frac:.4f} {relmse:.3g} mean {float(ref.mean()):.4f}")
        scene.close()
    print(json.dumps({"scenes": n, "failed": fails}))


if __name__ == "__main__":
    main()
