#!/usr/bin/env python
"""Render a scene with the B200 path and with the unmodified reference (oracle/_ref
ref_harness, same per-(pixel,sample) PCG32 streams) and print how the two films agree.
Developer tool (GPU box): python tools/compare_render.py scenes/smoke.pbrt [--spp N] [--res WxH]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def agreement(gpu_rgb, ref_rgb):
    """Per-pixel statistics of two float RGB images."""
    d = np.abs(gpu_rgb - ref_rgb)
    scale = np.maximum(np.abs(ref_rgb), 1e-3)
    rel = (d / scale).max(axis=-1)
    mse = ((gpu_rgb - ref_rgb) ** 2).mean(axis=-1)
    relmse = float((mse / (ref_rgb.mean(axis=-1) ** 2 + 1e-2)).mean())
    return {
        "mean_gpu": float(gpu_rgb.mean()), "mean_ref": float(ref_rgb.mean()),
        "relMSE": relmse,
        "frac_pixels_rel<1e-5": float((rel < 1e-5).mean()),
        "frac_pixels_rel<1e-3": float((rel < 1e-3).mean()),
        "frac_pixels_rel<1e-1": float((rel < 1e-1).mean()),
        "max_abs": float(d.max()),
    }


def main():
    from conftest import load_product
    import oracle
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--res", default="")
    ap.add_argument("--maxdepth", type=int, default=-1)
    ap.add_argument("--out", default="/tmp/iile_cmp")  # holds the flat scene: keep it out of gpurun_out/
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    iile = load_product()
    out = Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    scene = iile.Scene.from_pbrt(a.scene, upload=True)
    setup = scene.info().setup
    if a.spp:
        setup.pixelSamples = a.spp
    if a.res:
        setup.xRes, setup.yRes = map(int, a.res.split("x"))
    if a.maxdepth >= 0:
        setup.maxDepth = a.maxdepth
    scene.set_setup(setup)
    flat = out / "scene.flat"
    scene.save_flat(flat)
    scene.render()                      # warm-up: lazy module load, path-state allocation
    t0 = time.time()
    film, stats = scene.render()
    wall = time.time() - t0
    rgb = scene.resolve(film)
    res = {"gpu": {"wall_s": wall, "device_s": stats.deviceSeconds, "cameraRays": stats.cameraRays,
                   "closestRays": stats.closestRays, "shadowRays": stats.shadowRays, "launches": stats.launches,
                   "msamples_per_s": stats.cameraRays / max(stats.deviceSeconds, 1e-9) / 1e6}}
    np.save(out / "gpu_rgb.npy", rgb)
    if not a.no_ref:
        info = oracle.ref_render(flat, out / "ref.pfm", sampler="pixelstream")
        ref = oracle.read_pfm(out / "ref.pfm")
        res["ref"] = info
        res["agreement"] = agreement(rgb, ref)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
