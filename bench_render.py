"""bench.py's path-traced leg: Msamples/s of SamplerIntegrator::Render on the B200.

Workload: BASELINE configs[0] — killeroo-simple, 400x400, `Integrator "path" maxdepth 5`, random
sampler — at 16 spp per GPU (the spp BASELINE names) and at 256 spp per GPU (steady state: 16
batches of 2.56 M paths).  Every rank renders its own sample range [rank*spp, (rank+1)*spp) of
every pixel into a full-frame {R,G,B,weight} film (SURVEY §8e partition A, "weak": per-GPU work
is fixed), then the films are summed onto rank 0 with one NCCL reduce — the MergeFilmTile step
(film.cpp:135-148) across GPUs — inside the timed region.
"""
from __future__ import annotations

import os
import time

import numpy as np


def run(iile, oracle, args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    scene_file = oracle.REF_SCENES / "cfg1.pbrt"
    if not scene_file.exists():
        return None
    scene = iile.Scene.from_pbrt(scene_file, upload=True)
    base = scene.info().setup
    h, w = scene.film_shape()
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    film = torch.zeros(h, w, 4, device="cuda")
    out = {}
    for spp in (16, 256):
        setup = scene.info().setup
        setup.pixelSamples = spp * world            # one image of spp*world samples per pixel
        scene.set_setup(setup)
        params = iile.RenderParams(rank * spp, (rank + 1) * spp, 0, 0, 1, 0, 0)

        def step():
            scene.render_dev(film.data_ptr(), params, st, want_stats=False)
            if world > 1:
                dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)

        steps = max(2, min(args.steps, 10 if spp == 16 else 3))
        for _ in range(3 if spp == 16 else 1):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        stats = scene.render_dev(film.data_ptr(), params, st, want_stats=True)   # ray counters of one pass
        rays = torch.tensor([stats.closestRays + stats.shadowRays, stats.cameraRays], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
            dist.all_reduce(rays)
        torch.cuda.synchronize()
        # end to end through the host-buffer call: film read back to pinned host memory every step
        t0 = time.perf_counter()
        host_film, _ = scene.render(params)
        e2e_s = time.perf_counter() - t0
        entry = {
            "msamples_per_s": float(rays[1].item()) / (ms * 1e-3) / 1e6,
            "mrays_per_s": float(rays[0].item()) / (ms * 1e-3) / 1e6,
            "ms_per_step": ms, "spp_per_gpu": spp, "samples": int(rays[1].item()), "rays": int(rays[0].item()),
            "launches_per_step": int(stats.launches),
            "e2e_msamples_per_s_rank0": float(stats.cameraRays) / e2e_s / 1e6,
            "film_reduce": "NCCL reduce of %d bytes per rank" % (h * w * 16) if world > 1 else "none (1 GPU)",
        }
        if rank == 0:
            rgb = scene.resolve(film.cpu().numpy())
            entry["film_mean"] = float(rgb.mean())
        out[f"cfg1_{spp}spp"] = entry
    scene.set_setup(base)
    if rank != 0:
        return None
    # CPU arm beside it: the unmodified reference on this box's cores (16 spp, bounded: ~1 s)
    if oracle.have_ref():
        tmp = f"/tmp/iile_bench_render_{os.getpid()}"
        os.makedirs(tmp, exist_ok=True)
        scene.save_flat(f"{tmp}/cfg1.flat")
        info = oracle.ref_render(f"{tmp}/cfg1.flat", f"{tmp}/ref.pfm", sampler="stock")
        out["cpu_reference"] = {"msamples_per_s": info["msamples_per_s"], "threads": info["threads"],
                                "sample": "cfg1 400x400x16spp, stock RandomSampler, PathIntegrator::Render wall time",
                                "film_mean": float(oracle.read_pfm(f"{tmp}/ref.pfm").mean())}
    out["workload"] = "cfg1: killeroo-simple 400x400, path maxdepth 5, random sampler (BASELINE configs[0])"
    return out
