#!/usr/bin/env python
"""Developer tool: closest-hit / any-hit kernel time as a function of the batch size (CUDA events)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product
import oracle, torch
iile = load_product()
scene = iile.Scene.from_pbrt(oracle.REF_SCENES / "cfg2.pbrt")
scene.save_flat("/tmp/t.flat")
flat = oracle.FlatScene("/tmp/t.flat")
lo, hi = oracle.killeroo_ray_box(flat)
N = 1 << 22
rays = torch.empty(N * 8, dtype=torch.float32, device="cuda")
hits = torch.empty(N * 4, dtype=torch.int32, device="cuda")
occ = torch.empty(N, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
iile.generate_rays_dev(lo, hi, 0, N, rays.data_ptr(), st)
torch.cuda.synchronize()
for n in (1, 1000, 10_000, 100_000, 300_000, 1_000_000, 1 << 22):
    for kind in ("closest", "any"):
        f = (lambda: scene.intersect_dev(rays.data_ptr(), n, hits.data_ptr(), st)) if kind == "closest" else \
            (lambda: scene.intersect_p_dev(rays.data_ptr(), n, occ.data_ptr(), st))
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); 
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        print(f"n={n:8d} {kind:8s} {e0.elapsed_time(e1)/10*1e3:9.1f} us")
