#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs[1], measured on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload = "cfg2"): the BVH ray-cast microbenchmark — 2^24 incoherent
random rays (SURVEY §8d.2 generator) closest-hit + the same 2^24 rays any-hit against the
killeroo geometry's SAH BVH (66,532 triangles, 118,375 LinearBVHNodes).  One step = both
casts over one batch, i.e. 2^25 ray casts per GPU.  Rays for N>1 are sharded by range (each
rank casts its own 2^24-ray range; no data-path collective; "weak" scaling).

  value     Mrays/s, whole job, rays resident in HBM, CUDA events on the launch stream,
            max over ranks
  e2e       same metric through the public host-buffer C-ABI calls (iile_intersect_closest /
            iile_intersect_any) with pinned host arrays: H2D of rays + D2H of hit records
            inside the timed region
  roofline  closest-hit kernel: algorithmic bytes (32 B x nodes fetched + 48 B x primitive
            tests + 32 B ray + 16 B hit, counts from the instrumented oracle on a sample of
            the same rays) / CUDA-event kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own BVHAccel (oracle/_ref, kind "reference") on all host cores
            over a bounded sample of the same rays
  also      Msamples/s of the path-traced cfg1 render (killeroo-simple 400x400x16spp) when the
            render path is available — the other half of BASELINE.json's metric string

`--impl reference` times the reference CPU implementation (oracle/_ref/ref_harness raycast)
on a bounded sample per step, on rank 0 only.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
N_RAYS = 1 << 24
METRIC = "Mrays/s BVH cast (closest-hit + any-hit), cfg2"
UNIT = "Mrays/s"


def load_product():
    name = "pbrt_v3_iile_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, ROOT / "pbrt-v3-iile_b200" / "__init__.py",
                                                  submodule_search_locations=[str(ROOT / "pbrt-v3-iile_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_oracle():
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    return oracle


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The region is tens of
    milliseconds, so NVML is polled in-process (~1 ms per sample) instead of waiting for
    `nvidia-smi -lms`; the fields are the ones the profiling recipe's nvidia-smi line reports."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.mask = [], 0
        self.stop_flag = threading.Event()
        self.thread = None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "NVML unavailable"}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        reasons = sorted(k for k, bit in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "reasons": reasons, "samples": len(self.sm), "source": "NVML polled during the timed region"}


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
    capture of this kernel on this workload (profiles/traffic.json), or None."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None
    t = json.loads(p.read_text()).get(kernel)
    return None if t is None else float(t["dram_bytes_read"] + t["dram_bytes_write"])


def cfg2_scene_path(oracle):
    p = oracle.REF_SCENES / "cfg2.pbrt"
    if not p.exists():
        raise SystemExit("bench.py: oracle/_ref/scenes/cfg2.pbrt is missing — run `make -C oracle ref` "
                         "once where /root/reference exists (the build products travel with the repo)")
    return p


def terrain_scene_path(rank):
    """BASELINE configs[4] geometry: the seeded 8,000,000-triangle height field (tools/make_scenes.py
    cfg5), generated on this box (152 MB PLY, ~2 s) — the working set that does NOT fit in L2."""
    d = Path(f"/tmp/iile_cfg5_{rank}")
    if not (d / "cfg5.pbrt").exists():
        subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_scenes.py"), "cfg5", str(d)],
                              stdout=subprocess.DEVNULL)
    return d / "cfg5.pbrt"


def workload_scene(args, oracle, rank):
    """(scene file, ray-origin box function, name).  Rays follow SURVEY §8d.2 for both."""
    if args.workload == "terrain":
        return terrain_scene_path(rank), oracle.cfg2_ray_box, "cfg2 ray generator on the cfg5 terrain (8M triangles)"
    return cfg2_scene_path(oracle), oracle.killeroo_ray_box, "cfg2"


# ----------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, all host cores, bounded sample."""
    if rank != 0:
        return
    oracle = load_oracle()
    iile = load_product()
    if not oracle.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return
    tmp = Path("/tmp/iile_bench_ref")
    tmp.mkdir(exist_ok=True)
    scene_file, ray_box, wl_name = workload_scene(args, oracle, 0)
    scene = iile.Scene.from_pbrt(scene_file, upload=False)
    scene.save_flat(tmp / "cfg2.flat")
    flat = oracle.FlatScene(tmp / "cfg2.flat")
    lo, hi = ray_box(flat)
    sample = 1 << 22
    rays = oracle.generate_rays(lo, hi, 0, sample)
    cores = os.cpu_count() or 1
    times = []
    for it in range(args.warmup + args.steps):
        _, a = oracle.ref_raycast(tmp / "cfg2.flat", rays, tmp, any_hit=False)
        _, b = oracle.ref_raycast(tmp / "cfg2.flat", rays, tmp, any_hit=True)
        if it >= args.warmup:
            times.append(a["seconds"] + b["seconds"])
    t = float(np.mean(times))
    value = 2 * sample / t / 1e6
    sample_desc = f"first 2^22 of the 2^24 cfg2 rays, closest-hit + any-hit, per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "rays_per_kind": sample, "triangles": int((flat.prims["type"] == 0).sum()),
                   "nodes": int(len(flat.nodes)), "bvh": "SAH maxnodeprims 4"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    iile = load_product()
    oracle = None
    if iile.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    iile.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as oracle_mod          # scene path + (rank 0) cpu_baseline / roofline counts only
    oracle = oracle_mod
    scene_file, ray_box, wl_name = workload_scene(args, oracle, rank)
    scene = iile.Scene.from_pbrt(scene_file, upload=True)
    info = scene.info()
    tmp = Path(f"/tmp/iile_bench_{rank}")
    tmp.mkdir(exist_ok=True)
    scene.save_flat(tmp / "cfg2.flat")
    flat = oracle.FlatScene(tmp / "cfg2.flat")
    lo, hi = ray_box(flat)

    n = N_RAYS
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
    hits = torch.empty(n * 4, dtype=torch.int32, device="cuda")
    occ = torch.empty(n, dtype=torch.uint8, device="cuda")
    iile.generate_rays_dev(lo, hi, rank * n, n, rays.data_ptr(), st)
    torch.cuda.synchronize()

    def step():
        scene.intersect_dev(rays.data_ptr(), n, hits.data_ptr(), st)
        scene.intersect_p_dev(rays.data_ptr(), n, occ.data_ptr(), st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    # ---- device-resident timing (value) + per-kernel events (roofline) -------------------
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        ev[3 * k].record(stream)
        scene.intersect_dev(rays.data_ptr(), n, hits.data_ptr(), st)
        ev[3 * k + 1].record(stream)
        scene.intersect_p_dev(rays.data_ptr(), n, occ.data_ptr(), st)
        ev[3 * k + 2].record(stream)
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    ms_closest = float(np.mean([ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(args.steps)]))
    ms_any = float(np.mean([ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(args.steps)]))
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * 2 * n / (ms_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI calls (pinned host memory) ---------------
    h_rays = torch.empty(n * 8, dtype=torch.float32, pin_memory=True)
    h_rays.copy_(rays)
    h_hits = torch.empty(n * 4, dtype=torch.int32, pin_memory=True)
    h_occ = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    C = iile.C

    def e2e_step():
        iile._check(iile.lib.iile_intersect_closest(scene._h, C.c_void_p(h_rays.data_ptr()), n, C.c_void_p(h_hits.data_ptr())))
        iile._check(iile.lib.iile_intersect_any(scene._h, C.c_void_p(h_rays.data_ptr()), n, C.c_void_p(h_occ.data_ptr())))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * 2 * n / float(e2e_s.item()) / 1e6
    # result sanity: the timed kernels did the work
    nhit = int((hits.view(n, 4)[:, 0] >= 0).sum())
    assert int(occ.sum()) == nhit, "any-hit and closest-hit disagree"

    also = None
    if not args.no_render and args.workload == "cfg2":
        also = render_leg(iile, oracle, args, rank, world, local_rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline: algorithmic bytes from the instrumented oracle on a sample of THESE rays --
    sample = 1 << 20
    r_s = rays.view(n, 8)[:sample].cpu().numpy().view(iile.RAY_DTYPE).reshape(-1)
    want, nodes_c, tests_c = oracle.cast(flat, r_s)
    got = hits.view(n, 4)[:sample].cpu().numpy().view(iile.HIT_DTYPE).reshape(-1)
    parity = bool(np.array_equal(got["prim"], want["prim"]) and np.array_equal(got["t"].view("u4"), want["t"].view("u4")))
    _, nodes_a, tests_a = oracle.cast(flat, r_s, any_hit=True)
    bytes_closest = (32 * nodes_c + 48 * tests_c) / sample + 48
    bytes_any = (32 * nodes_a + 48 * tests_a) / sample + 48
    peak, peak_src = measured_peak()
    achieved = n * bytes_closest / (ms_closest * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "closestHitKernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": measured_traffic("closestHitKernel" if args.workload == "cfg2" else "closestHitKernel@terrain"),
                "algorithmic_bytes_per_launch": n * bytes_closest, "peak_source": peak_src,
                "bytes_per_ray": bytes_closest, "ms_per_launch": ms_closest,
                "any_hit": {"bytes_per_ray": bytes_any, "ms_per_launch": ms_any,
                            "achieved": n * bytes_any / (ms_any * 1e-3) / 1e9},
                "note": ("working set (3.8 MB nodes + 3.2 MB triangles) is L2-resident: latency/issue-bound, "
                         "not an HBM test; traffic from ncu in profiles/") if args.workload == "cfg2" else
                        ("working set ~%.0f MB of 4-wide nodes + %.0f MB of triangle slots: larger than the 126 MB L2"
                         % (info.nNodes / 4 * 128 / 1e6, info.nPrims * 48 / 1e6))}
    # ---- CPU baseline: the reference's BVHAccel on this box's cores, bounded sample ---------
    cpu = None
    if args.no_cpu_baseline:
        cpu = None
    elif oracle.have_ref():
        cs = 1 << 22
        r_c = rays.view(n, 8)[:cs].cpu().numpy().view(iile.RAY_DTYPE).reshape(-1)
        _, a = oracle.ref_raycast(tmp / "cfg2.flat", r_c, tmp, any_hit=False)
        _, b = oracle.ref_raycast(tmp / "cfg2.flat", r_c, tmp, any_hit=True)
        cpu = {"value": 2 * cs / (a["seconds"] + b["seconds"]) / 1e6, "unit": UNIT, "cores": os.cpu_count(),
               "kind": "reference", "sample": "first 2^22 of rank 0's cfg2 rays, closest-hit + any-hit, "
               "oracle/_ref BVHAccel::Intersect/IntersectP via ParallelFor",
               "closest_mrays_s": a["mrays_per_s"], "any_mrays_s": b["mrays_per_s"]}
    else:
        t0 = time.perf_counter()
        oracle.cast(flat, r_s)
        oracle.cast(flat, r_s, any_hit=True)
        cpu = {"value": 2 * sample / (time.perf_counter() - t0) / 1e6, "unit": UNIT, "cores": os.cpu_count(),
               "kind": "port", "sample": "first 2^20 rays, oracle restatement, all cores"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "rays_per_kind_per_gpu": n, "kinds": ["closest-hit", "any-hit"],
                   "triangles": int(info.nTriangles), "nodes": int(info.nNodes), "bvh": "SAH maxnodeprims 4",
                   "l2_policy": "inputs larger than L2 (512 MiB of rays per cast)",
                   "hit_fraction": nhit / n, "parity_vs_oracle_on_sample": parity},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 32,
                "d2h_bytes_per_step": n * 16 + n},
        "gpu_launches": 2 * args.steps,
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
    }
    if also:
        out["also"] = also
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def render_leg(iile, oracle, args, rank, world, local_rank):
    """Msamples/s path-traced (the other half of the metric string): cfg1, sample-range
    partitioned across ranks, film reduced with NCCL.  Returns None until iile_render exists."""
    try:
        import bench_render
    except Exception:
        return None
    return bench_render.run(iile, oracle, args, rank, world, local_rank)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-render", action="store_true", help="skip the path-traced 'also' leg")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "terrain"],
                    help="cfg2 = BASELINE configs[1] (default, the bench line); terrain = the same ray "
                         "generator on the 8M-triangle cfg5 geometry (HBM-resident working set)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
