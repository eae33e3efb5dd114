"""bench.py's path-traced leg, and a stand-alone bench for the rendered BASELINE configs.

As bench.py's "also" leg (`run`): BASELINE configs[0] — killeroo-simple, 400x400, `Integrator "path"
maxdepth 5`, random sampler — at 16 spp per GPU (the spp BASELINE names) and at 256 spp per GPU
(steady state: 16 batches of 2.56 M paths).

Stand-alone (`python bench_render.py --workload cfg1|cfg3|cfg5 [--spp S] [--steps K]`, or under
torchrun for N GPUs): one JSON line, metric Msamples/s path-traced, for
  cfg3  configs[2]: generated room, 16 area lights (spatial light distribution), 798k triangles,
        1024x1024, maxdepth 8 (default 32 spp per GPU per step; BASELINE's 256 with --spp 256)
  cfg5  configs[4]: 8,000,000-triangle terrain, 3840x2160, maxdepth 5 (default 8 spp per GPU per
        step; BASELINE's 64 spp = --spp 64 at N=1, 8 at N=8)
  cfg4  configs[3]: the IISPT hemispheric auxiliary pass — one batch of 4096 32x32 hemispheres
        (1 spp, maxdepth 3: intensity + camera-space normals + distance) per GPU per step at hit
        points of the cfg1 scene, through iile_render_hemispheres (host buffers: the call
        IisptRenderRunner would make); metric hemispheres/s, ranks take disjoint batches

Every rank renders its own sample range [rank*spp, (rank+1)*spp) of every pixel into a full-frame
{R,G,B,weight} film (SURVEY §8e partition A, "weak": per-GPU work is fixed), then the films are
summed onto rank 0 with one NCCL reduce — the MergeFilmTile step (film.cpp:135-148) across GPUs —
inside the timed region.
"""
from __future__ import annotations

import os
import time

import numpy as np


_COMM = {}


def film_comm(iile, rank, world):
    """One FilmComm per process: the NCCL communicator behind iile_film_reduce_dev (the unique id
    travels over the torch.distributed group the launcher set up)."""
    if world > 1 and "c" not in _COMM:
        _COMM["c"] = iile.partition.film_comm(rank, world)
    return _COMM.get("c")


def time_render(iile, scene, spp, steps, warm, rank, world):
    """Device-timed render of `spp` samples per pixel per GPU (+ NCCL film reduce), max over ranks."""
    import torch
    import torch.distributed as dist

    h, w = scene.film_shape()
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    film = torch.zeros(h, w, 4, device="cuda")
    setup = scene.info().setup
    setup.pixelSamples = spp * world            # one image of spp*world samples per pixel
    scene.set_setup(setup)
    params = iile.RenderParams(rank * spp, (rank + 1) * spp, 0, 0, 1, 0, 0)
    comm = film_comm(iile, rank, world)

    def step():
        scene.render_dev(film.data_ptr(), params, st, want_stats=False)
        if world > 1:   # Film::MergeFilmTile across GPUs: iile_film_reduce_dev on the render's stream
            comm.reduce(film.data_ptr(), h * w, 0, st)

    for _ in range(warm):
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
    rays = torch.tensor([stats.closestRays + stats.shadowRays, stats.cameraRays, stats.closestRays, stats.shadowRays],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        comm.reduce(film.data_ptr(), h * w, 0, st)
        dist.all_reduce(rays)
    torch.cuda.synchronize()
    # end to end through the host-buffer call (iile_render): the film is read back into the caller's
    # page-locked host buffer every step; mean of `steps` calls after one warm-up
    pinned = torch.empty(h, w, 4, dtype=torch.float32, pin_memory=True)
    host_view = pinned.numpy()
    host_film, _ = scene.render(params, out=host_view)
    t0 = time.perf_counter()
    for _ in range(max(1, steps)):
        host_film, _ = scene.render(params, out=host_view)
    e2e_s = (time.perf_counter() - t0) / max(1, steps)
    entry = {
        "msamples_per_s": float(rays[1].item()) / (ms * 1e-3) / 1e6,
        "mrays_per_s": float(rays[0].item()) / (ms * 1e-3) / 1e6,
        "ms_per_step": ms, "spp_per_gpu": spp, "samples": int(rays[1].item()), "rays": int(rays[0].item()),
        "closest_rays": int(rays[2].item()), "shadow_rays": int(rays[3].item()),
        "launches_per_step": int(stats.launches),
        "e2e_msamples_per_s_rank0": float(stats.cameraRays) / e2e_s / 1e6,
        "film_reduce": "iile_film_reduce_dev (ncclReduce, NCCL %d) of %d bytes per rank" % (iile.FilmComm.nccl_version(), h * w * 16) if world > 1 else "none (1 GPU)",
    }
    if rank == 0:
        rgb = scene.resolve(film.cpu().numpy())
        entry["film_mean"] = float(rgb.mean())
    return entry


def run(iile, oracle, args, rank, world, local_rank):
    scene_file = oracle.REF_SCENES / "cfg1.pbrt"
    if not scene_file.exists():
        return None
    scene = iile.Scene.from_pbrt(scene_file, upload=True)
    base = scene.info().setup
    out = {}
    for spp in (16, 256):
        steps = max(2, min(args.steps, 10 if spp == 16 else 3))
        out[f"cfg1_{spp}spp"] = time_render(iile, scene, spp, steps, 3 if spp == 16 else 1, rank, world)
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



def hemisphere_points(iile, oracle, scene, flat_path, n, first_ray):
    """Hemisphere origins + axes: hit points of cfg2-style random rays, pushed off
    the surface along the geometric normal facing the ray."""
    flat = oracle.FlatScene(flat_path)
    lo, hi = oracle.killeroo_ray_box(flat)
    pos, nrm = [], []
    have = 0
    while have < n:
        rays = oracle.generate_rays(lo, hi, first_ray, 4 * n)
        first_ray += 4 * n
        hits = scene.intersect(rays)
        ok = np.flatnonzero((hits["prim"] >= 0) & (flat.prims["type"][np.maximum(hits["prim"], 0)] == 0))
        pr = flat.prims[hits["prim"][ok]]
        p0, p1, p2 = (flat.P[pr["v"][:, k]] for k in range(3))
        ng = np.cross(p0 - p2, p1 - p2)
        ng /= np.maximum(np.linalg.norm(ng, axis=1, keepdims=True), 1e-20)
        d = rays["d"][ok]
        ng = np.where((ng * d).sum(axis=1, keepdims=True) > 0, -ng, ng).astype(np.float32)
        p = (rays["o"][ok] + hits["t"][ok, None] * d + np.float32(0.05) * ng).astype(np.float32)
        pos.append(p)
        nrm.append(ng)
        have += len(p)
    return np.concatenate(pos)[:n], np.concatenate(nrm)[:n]


def run_cfg4(iile, oracle, args, rank, world):
    import torch
    import torch.distributed as dist
    import json
    from pathlib import Path
    import bench

    n = args.spp or 4096            # hemispheres per GPU per step
    scene = iile.Scene.from_pbrt(oracle.REF_SCENES / "cfg1.pbrt", upload=True)
    tmp = Path(f"/tmp/iile_bench_hemi_{rank}")
    tmp.mkdir(exist_ok=True)
    scene.save_flat(tmp / "cfg1.flat")
    pos, axis = hemisphere_points(iile, oracle, scene, tmp / "cfg1.flat", n, rank * (1 << 24))
    # the caller's page-locked result buffers (what a host integrating the path would own)
    bufs = tuple(torch.empty(*shape, dtype=torch.float32, pin_memory=True).numpy()
                 for shape in ((n, 32, 32, 3), (n, 32, 32, 3), (n, 32, 32)))
    for _ in range(max(args.warmup, 1)):
        scene.render_hemispheres(pos, axis, out=bufs)
    clocks = bench.ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    t0 = time.perf_counter()
    dev_s, rays = 0.0, 0
    for _ in range(args.steps):
        inten, nrm, dist_, stats = scene.render_hemispheres(pos, axis, out=bufs)
        dev_s += stats.deviceSeconds
        rays = stats.closestRays + stats.shadowRays
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps
    clk = clocks.stop()
    t = torch.tensor([wall, dev_s / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev = float(t[0]), float(t[1])
    if rank != 0:
        return
    out = {"metric": "hemispheres/s, IISPT hemispheric pass (32x32, 1 spp, intensity+normals+distance), cfg4",
           "value": world * n / dev, "unit": "hemispheres/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 1), "ms_per_step": dev * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "cfg4", "scene": "cfg1 (killeroo-simple geometry)", "hemispheres_per_gpu_per_step": n,
                      "size": 32, "spp": 1, "maxdepth": 3},
           "e2e": {"value": world * n / wall, "unit": "hemispheres/s", "h2d_bytes_per_step": n * 32 * 4,
                   "d2h_bytes_per_step": n * 1024 * (12 + 12 + 4),
                   "note": "iile_render_hemispheres with host buffers: camera upload, kernels, film/normal/distance read-back"},
           "msamples_per_s": world * n * 1024 / dev / 1e6, "mrays_per_s": world * rays / dev / 1e6,
           "distance_hit_fraction": float((dist_ >= 0).mean()), "intensity_mean": float(inten.mean()),
           "clocks": clk}
    # the same batch left in HBM as the network's input (iile_hemispheres_to_nn_dev: RenderView +
    # normalizeMapsDownstream on the device) and the inverse transform of a network-sized answer
    nn = torch.empty(n, 7, 32, 32, device="cuda")
    means = torch.empty(n, 3, device="cuda")
    back = torch.empty(n, 32, 32, 3, device="cuda")
    stream = torch.cuda.current_stream()
    scene.hemispheres_to_nn_dev(pos, axis, nn.data_ptr(), means.data_ptr(), stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    for _ in range(args.steps):
        scene.hemispheres_to_nn_dev(pos, axis, nn.data_ptr(), means.data_ptr(), stream=stream.cuda_stream)
    e1.record(stream)
    for _ in range(args.steps):
        iile.hemispheres_from_nn_dev(n, 32, nn.data_ptr(), means.data_ptr(), back.data_ptr(), stream.cuda_stream)
    e2.record(stream)
    torch.cuda.synchronize()
    to_ms, from_ms = e0.elapsed_time(e1) / args.steps, e1.elapsed_time(e2) / args.steps
    out["nn_path"] = {"to_nn_hemispheres_per_s": n / (to_ms * 1e-3), "to_nn_ms_per_step": to_ms,
                      "from_nn_ms_per_step": from_ms, "nn_input_bytes": n * 7 * 1024 * 4,
                      "note": "rank 0, device-timed: iile_hemispheres_to_nn_dev leaves [n,7,32,32] + means in HBM "
                              "(no read-back); iile_hemispheres_from_nn_dev = transformMapsUpstream of [n,3,32,32]"}
    if oracle.have_ref():
        k = 64
        _, _, _, info = oracle.ref_hemispheres(tmp / "cfg1.flat", pos[:k], axis[:k], tmp)
        out["cpu_baseline"] = {"value": info["hemispheres_per_s"], "unit": "hemispheres/s", "cores": os.cpu_count(),
                               "kind": "reference", "sample": f"first {k} hemispheres of the batch, IISPTdIntegrator::RenderView "
                               "one hemisphere after another (as IisptRenderRunner calls it), incl. its PFM writes"}
    print(json.dumps(out))


def render_roofline(iile, oracle, scene, entry, world):
    """SURVEY §8d for a rendered config: per sample, sum over its rays of 32 n_nodes + 48 n_tris + 48,
    plus 32 B of film.  The node / triangle counts are the reference traversal's (instrumented oracle)
    on 2^17 random rays that start INSIDE this scene's triangle bound (uniform origin, uniform direction:
    like path vertices they begin among the geometry), closest-hit and any-hit, weighted by the render's
    own closest / shadow ray counts.  An estimate: the path rays themselves are not replayed on the CPU."""
    import bench
    import numpy as np
    from pathlib import Path
    tmp = Path(f"/tmp/iile_bench_roof_{os.getpid()}")
    tmp.mkdir(exist_ok=True)
    scene.save_flat(tmp / "scene.flat")
    flat = oracle.FlatScene(tmp / "scene.flat")
    (tmp / "scene.flat").unlink()
    lo, hi = flat.triangle_bounds()
    n = 1 << 17
    rays = oracle.generate_rays(lo, hi, 0, n)
    _, nodes_c, tests_c = oracle.cast(flat, rays)
    _, nodes_a, tests_a = oracle.cast(flat, rays, any_hit=True)
    b_c = (32 * nodes_c + 48 * tests_c) / n + 48
    b_a = (32 * nodes_a + 48 * tests_a) / n + 48
    total = entry["closest_rays"] * b_c + entry["shadow_rays"] * b_a + entry["samples"] * 32
    peak, src = bench.measured_peak()
    achieved = total / (entry["ms_per_step"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "castKernel + shadeKernel (whole step)", "achieved": achieved,
            "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world), "traffic": None,
            "algorithmic_bytes_per_step": total, "bytes_per_closest_ray": b_c, "bytes_per_shadow_ray": b_a,
            "bytes_per_sample": total / max(entry["samples"], 1), "peak_source": src,
            "note": "estimate: per-ray node / triangle counts of the reference traversal (instrumented oracle) on 2^17 "
                    "random rays starting inside this scene's bound, weighted by the render's closest / shadow ray "
                    "counts (SURVEY 8d formula); queue and path-state traffic deliberately not credited"}


def main():
    """Stand-alone: one JSON line for a rendered BASELINE config (see the module docstring)."""
    import argparse
    import json
    import subprocess
    import sys
    from pathlib import Path

    import torch
    import torch.distributed as dist
    import bench

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg3", "cfg4", "cfg5", "cfg1d", "cfg3d"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per GPU per step (cfg4: hemispheres per GPU per step)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-spp", type=int, default=0, help="also time the reference on this box at this spp (rank 0)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    iile = bench.load_product()
    oracle = bench.load_oracle()
    torch.cuda.set_device(local_rank)
    iile.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload == "cfg4":
        run_cfg4(iile, oracle, args, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "cfg5":
        scene_file = bench.terrain_scene_path(0) if world > 1 and rank == 0 else (
            bench.terrain_scene_path(rank) if world == 1 else None)
    else:
        scene_file = oracle.REF_SCENES / f"{args.workload}.pbrt"
    spp = args.spp or {"cfg1": 16, "cfg3": 32, "cfg5": 8, "cfg1d": 16, "cfg3d": 32}[args.workload]
    # one parse + BVH build per node: rank 0 flattens, the others load its flat scene
    scene = iile.partition.load_scene_shared(iile.Scene, scene_file, rank, world, upload=True)
    info = scene.info()
    clocks = bench.ClockSampler(local_rank)
    clocks.start()
    entry = time_render(iile, scene, spp, args.steps, max(args.warmup, 1), rank, world)
    clk = clocks.stop()
    if rank == 0:
        out = {"metric": "Msamples/s path-traced, " + args.workload, "value": entry["msamples_per_s"],
               "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
               "ms_per_step": entry["ms_per_step"], "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": args.workload, "resolution": [int(info.setup.xRes), int(info.setup.yRes)],
                          "spp_per_gpu_per_step": spp, "maxdepth": int(info.setup.maxDepth),
                          "triangles": int(info.nTriangles), "nodes": int(info.nNodes), "lights": int(info.nLights),
                          "partition": "sample range per rank, scene replicated, NCCL film reduce",
                          "bvh_build_s": float(info.buildSeconds)},
               "clocks": clk, "detail": entry}
        try:
            out["roofline"] = render_roofline(iile, oracle, scene, entry, world)
        except Exception as ex:   # the roofline is commentary: never lose the measurement over it
            out["roofline"] = {"error": str(ex)}
        if args.cpu_spp and oracle.have_ref():
            tmp = Path(f"/tmp/iile_bench_render_{os.getpid()}")
            tmp.mkdir(exist_ok=True)
            setup = scene.info().setup
            setup.pixelSamples = args.cpu_spp
            scene.set_setup(setup)
            scene.save_flat(tmp / "scene.flat")
            ref = oracle.ref_render(tmp / "scene.flat", tmp / "ref.pfm", sampler="stock")
            out["cpu_baseline"] = {"value": ref["msamples_per_s"], "unit": "Msamples/s", "cores": ref["threads"],
                                   "kind": "reference",
                                   "sample": f"{args.workload} at {args.cpu_spp} spp, stock sampler, "
                                             "PathIntegrator::Render wall time (per-spp cost is constant)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
