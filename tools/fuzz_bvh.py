#!/usr/bin/env python
"""The four BVH builders (sah / middle / equal / hlbvh) on random and deliberately awkward triangle soups — duplicated
triangles, zero-area triangles, coincident centroids, axis-aligned sheets (flat bounds), wildly different scales, a few
spheres in between — against the reference's own BVHAccel built single-threaded (ref_harness bvh): LinearBVHNode bytes
(the reference's uninitialised leaf `axis` and pad bytes aside) and the primitive order.  Runs on the CPU; with
`device` as third argument the soups have 65,536+ primitives and splitmethod "hlbvh", i.e. the tree is built by
csrc/cuda/hlbvh_device.cu (needs a GPU).  python tools/fuzz_bvh.py [n] [seed] [device]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402


def soup(rng, n, kind):
    c = rng.uniform(-1, 1, (n, 3))
    e = rng.uniform(-0.1, 0.1, (n, 2, 3))
    if kind == "scales":
        e *= 10.0 ** rng.uniform(-4, 1, (n, 1, 1))
    if kind == "sheet":
        c[:, 2] = 0.25
        e[:, :, 2] = 0
    if kind == "lattice":  # many equal centroids along every axis
        c = np.round(c * 3) / 3
    tri = np.stack([c, c + e[:, 0], c + e[:, 1]], axis=1)
    if kind == "duplicates":
        tri[n // 2:] = tri[:n - n // 2]
    if kind == "degenerate":
        tri[::3, 2] = tri[::3, 1]          # zero-area
        tri[1::7] = tri[1::7, :1]          # a single point
    return tri.astype(np.float32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    device = len(sys.argv) > 3 and sys.argv[3] in ("device", "big")  # big: the same soups through the host builder
    on_gpu = device and sys.argv[3] == "device"
    iile = load_product()
    if on_gpu:
        iile.set_device(0)
    work = Path("/tmp/fuzz_bvh")
    work.mkdir(exist_ok=True)
    fails = 0
    for k in range(n):
        kind = ["plain", "scales", "sheet", "lattice", "duplicates", "degenerate"][k % 6]
        nt = int(rng.choice([1, 2, 3, 5, 17, 100, 1000, 5000]))
        split = ["sah", "middle", "equal", "hlbvh"][int(rng.integers(4))]
        if device:
            nt, split = int(rng.choice([65536, 65537, 90001, 250000])), "hlbvh"
        leaf = int(rng.choice([1, 2, 4, 7, 16, 255]))
        tri = soup(rng, nt, kind)
        P = " ".join(f"{v:.9g}" for v in tri.reshape(-1))
        idx = " ".join(str(i) for i in range(3 * nt))
        spheres = ""
        for _ in range(int(rng.integers(0, 3))):
            t = rng.uniform(-1, 1, 3)
            spheres += f'AttributeBegin\nTranslate {t[0]:.5f} {t[1]:.5f} {t[2]:.5f}\nShape "sphere" "float radius" [{rng.uniform(0.01, 0.4):.5f}]\nAttributeEnd\n'
        text = (f'LookAt 0 -5 0 0 0 0 0 0 1\nCamera "perspective"\nFilm "image" "integer xresolution" [8] "integer yresolution" [8]\n'
                f'Accelerator "bvh" "string splitmethod" "{split}" "integer maxnodeprims" [{leaf}]\nWorldBegin\n'
                f'LightSource "point"\nMaterial "matte"\n{spheres}Shape "trianglemesh" "integer indices" [{idx}] "point P" [{P}]\nWorldEnd\n')
        src = work / "s.pbrt"
        src.write_text(text)
        s = iile.Scene.from_pbrt(src, upload=on_gpu)
        flat = work / "s.flat"
        s.save_flat(flat)
        nodes, order, info = oracle.ref_bvh(flat, work)
        mine = s.nodes()
        ok = len(mine) == len(nodes) and np.array_equal(s.prim_order(), order)
        if ok:
            for name in ("bmin", "bmax", "offset", "nPrims"):
                ok = ok and np.ascontiguousarray(mine[name]).tobytes() == np.ascontiguousarray(nodes[name]).tobytes()
            interior = nodes["nPrims"] == 0
            ok = ok and np.array_equal(mine["axis"][interior], nodes["axis"][interior])
        if not ok:
            fails += 1
            (work / f"fail_{fails}.pbrt").write_text(text)
            detail = {"nodes": [len(mine), len(nodes)], "order_equal": bool(len(order) == len(s.prim_order()) and np.array_equal(s.prim_order(), order))}
            if len(mine) == len(nodes):
                for name in ("bmin", "bmax", "offset", "nPrims"):
                    a, b = mine[name], nodes[name]
                    bad = np.flatnonzero((a != b).reshape(len(a), -1).any(axis=1))
                    if len(bad):
                        detail[name] = [int(len(bad)), int(bad[0]), a[bad[0]].tolist(), b[bad[0]].tolist()]
            else:
                m = min(len(mine), len(nodes))
                bad = np.flatnonzero((mine["nPrims"][:m] != nodes["nPrims"][:m]) | (mine["offset"][:m] != nodes["offset"][:m]))
                detail["first_structural_difference"] = int(bad[0]) if len(bad) else m
                if len(bad):
                    k0 = int(bad[0])
                    detail["mine"] = [int(mine["nPrims"][k0]), int(mine["offset"][k0])]
                    detail["ref"] = [int(nodes["nPrims"][k0]), int(nodes["offset"][k0])]
            print("     ", json.dumps(detail))
        print("ok  " if ok else "FAIL", kind, nt, split, leaf, len(nodes))
        s.close()
    print(json.dumps({"trees": n, "failed": fails}))


if __name__ == "__main__":
    main()
