#!/usr/bin/env python
"""Wall time of one IisptRenderRunner::run task on the device (hemi points -> RenderView + normalise -> identity
network -> upstream -> gather) next to the reference's own members walking the same task on the host
(ref_harness iispt_task, one core, as IisptRenderRunner runs per thread).  python tools/time_iispt_task.py [size] [tile]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402


def main():
    import torch
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tile = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    iile = load_product()
    iile.set_device(0)
    scene = iile.Scene.from_pbrt(oracle.REF_SCENES / "cfg1.pbrt")
    x0 = y0 = (400 - size) // 2
    task = iile.IisptTask(x0, y0, x0 + size, y0 + size, tile, 32, 0, 0, 7)
    net = lambda x: x[:, 0:3]
    scene.iispt_task(task, net)                     # warm-up: buffers, light tables
    torch.cuda.synchronize()
    t = time.time()
    reps = 5
    for _ in range(reps):
        samples, pos, direction, valid, films = scene.iispt_task(task, net)
    torch.cuda.synchronize()
    gpu_s = (time.time() - t) / reps
    out = {"task": [x0, y0, x0 + size, y0 + size, tile], "hemi_points": int(len(valid)), "valid": int(valid.sum()),
           "pixels": size * size, "gpu_seconds_per_task": gpu_s,
           "gpu_hemispheres_per_s": float(valid.sum()) / gpu_s, "gpu_pixels_per_s": size * size / gpu_s}
    if oracle.have_ref():
        flat = "/tmp/_iispt_time.flat"
        scene.save_flat(flat)
        _, _, _, info = oracle.ref_iispt_task(flat, "/tmp", x0, y0, x0 + size, y0 + size, tile, 32, 0, 0, 7)
        out["reference_hemi_seconds"] = info["hemi_seconds"]
        out["reference_gather_seconds"] = info["gather_seconds"]
        out["speedup_vs_one_reference_thread"] = (info["hemi_seconds"] + info["gather_seconds"]) / gpu_s
    print(json.dumps(out))


if __name__ == "__main__":
    main()
