#!/usr/bin/env python
"""iile_intersect_closest / iile_intersect_any on awkward scenes (tools/fuzz_bvh.py's soups: duplicated triangles —
every ray through them is an exact tie —, zero-area triangles, flat sheets, lattices, six decades of scale, plus
spheres / disks / cylinders) with awkward rays (random, axis-aligned with zero direction components, origins ON
primitives' vertices and in sheet planes, short tMax) against the UNMODIFIED reference's BVHAccel::Intersect /
IntersectP on the same tree (ref_harness raycast): primitive and t bit-identical, b1 / b2 equal (the sign of a zero
is not observable through the reference's isect.uv); occlusion flags identical.
python tools/fuzz_cast.py [n] [seed]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tools"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402
from fuzz_bvh import soup  # noqa: E402


def rays_for(rng, tri, n):
    lo, hi = tri.reshape(-1, 3).min(0) - 0.5, tri.reshape(-1, 3).max(0) + 0.5
    r = np.zeros(n, dtype=oracle.RAY_DTYPE)
    o = rng.uniform(lo, hi, (n, 3))
    target = tri[rng.integers(len(tri), size=n), rng.integers(3, size=n)] + rng.normal(0, 0.02, (n, 3))
    d = target - o
    k = n // 6
    axis = rng.integers(3, size=k)                       # axis-aligned: two zero components
    d[:k] = 0
    d[np.arange(k), axis] = rng.choice([-1.0, 1.0], k)
    o[:k] = target[:k] - d[:k] * rng.uniform(0.1, 2, (k, 1))
    o[k:2 * k] = tri[rng.integers(len(tri), size=k), rng.integers(3, size=k)]   # origins on vertices
    d[2 * k:3 * k, rng.integers(3)] = 0                  # one zero component
    exact = target.copy()
    exact[3 * k:4 * k] = tri[rng.integers(len(tri), size=k), rng.integers(3, size=k)]  # aimed exactly at vertices
    d[3 * k:4 * k] = exact[3 * k:4 * k] - o[3 * k:4 * k]
    r["o"], r["d"] = o.astype(np.float32), d.astype(np.float32)
    r["tMax"] = np.inf
    r["tMax"][4 * k:5 * k] = rng.uniform(0.2, 1.2, k).astype(np.float32)
    bad = ~(np.abs(r["d"]).sum(axis=1) > 0)
    r["d"][bad] = (1, 0, 0)
    return r


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    iile = load_product()
    iile.set_device(0)
    work = Path("/tmp/fuzz_cast")
    work.mkdir(exist_ok=True)
    fails = 0
    for k in range(n):
        kind = ["plain", "scales", "sheet", "lattice", "duplicates", "degenerate"][k % 6]
        nt = int(rng.choice([1, 2, 5, 17, 100, 1000, 5000, 30000]))
        split = ["sah", "middle", "equal", "hlbvh"][int(rng.integers(4))]
        leaf = int(rng.choice([1, 2, 4, 16, 255]))
        tri = soup(rng, nt, kind)
        P = " ".join(f"{v:.9g}" for v in tri.reshape(-1))
        idx = " ".join(str(i) for i in range(3 * nt))
        quads = ""
        for _ in range(int(rng.integers(0, 4))):
            t = rng.uniform(-1, 1, 3)
            shape = ['Shape "sphere" "float radius" [%.5f]' % rng.uniform(0.05, 0.4),
                     'Shape "sphere" "float radius" [0.3] "float zmin" [-0.1] "float zmax" [0.2] "float phimax" [200]',
                     'Shape "disk" "float radius" [%.5f] "float innerradius" [0.05]' % rng.uniform(0.1, 0.5),
                     'Shape "cylinder" "float radius" [0.2] "float zmin" [-0.3] "float zmax" [0.3] "float phimax" [270]'][int(rng.integers(4))]
            quads += f'AttributeBegin\nTranslate {t[0]:.5f} {t[1]:.5f} {t[2]:.5f}\nRotate {rng.uniform(0, 360):.3f} 1 1 0\n{shape}\nAttributeEnd\n'
        text = (f'LookAt 0 -5 0 0 0 0 0 0 1\nCamera "perspective"\nFilm "image" "integer xresolution" [8] "integer yresolution" [8]\n'
                f'Accelerator "bvh" "string splitmethod" "{split}" "integer maxnodeprims" [{leaf}]\nWorldBegin\n'
                f'LightSource "point"\nMaterial "matte"\n{quads}Shape "trianglemesh" "integer indices" [{idx}] "point P" [{P}]\nWorldEnd\n')
        src = work / "s.pbrt"
        src.write_text(text)
        s = iile.Scene.from_pbrt(src)
        flat = work / "s.flat"
        s.save_flat(flat)
        rays = rays_for(rng, tri, 30000)
        ref, _ = oracle.ref_raycast(flat, rays, work)
        hits = s.intersect(rays)
        diff = (hits["prim"] != ref["prim"]) | (hits["t"].view("u4") != ref["t"].view("u4"))
        hit = ref["prim"] >= 0
        types = oracle.FlatScene(flat).prims["type"]
        tri_hit = hit & (types[np.maximum(ref["prim"], 0)] == 0)  # barycentrics: triangles only
        # the harness reads the barycentrics out of isect.uv = b0 * (0,0) + b1 * (1,0) + b2 * (0,1): the additions turn
        # a -0 into +0, so the comparison is by value (any other difference is at least an ulp)
        diff |= tri_hit & ((hits["b1"] != ref["b1"]) | (hits["b2"] != ref["b2"]))
        occ_ref, _ = oracle.ref_raycast(flat, rays, work, any_hit=True)
        occ = s.intersect_p(rays)
        nd, no = int(diff.sum()), int((occ != occ_ref).sum())
        ok = nd == 0 and no == 0
        if not ok:
            fails += 1
            (work / f"fail_{fails}.pbrt").write_text(text)
            rays[diff | (occ != occ_ref)][:64].tofile(work / f"fail_{fails}_rays.bin")
            j = int(np.flatnonzero(diff | (occ != occ_ref))[0])
            print("     first:", rays[j], "mine", hits[j], "ref", (int(ref["prim"][j]), float(ref["t"][j])), "occ", int(occ[j]), int(occ_ref[j]))
        print("ok  " if ok else "FAIL", kind, nt, split, leaf, f"hits {int(hit.sum())} of {len(rays)}, closest diffs {nd}, any diffs {no}")
        s.close()
    print(json.dumps({"scenes": n, "rays_per_scene": 30000, "failed": fails}))


if __name__ == "__main__":
    main()
