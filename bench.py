#!/usr/bin/env python
"""bench.py — BASELINE.json's metric ("Msamples/s path-traced + Mrays/s BVH cast, 1/2/4/8 B200 vs pbrt
CPU") measured on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload W] [--legs a,b]

Headline (config.workload = "cfg5", BASELINE configs[4], the one config that names the multi-GPU split):
the synthetic 8,000,000-triangle terrain, 3840x2160, Integrator "path" maxdepth 5, random sampler, 64 spp.
One step = the whole 64-spp image.  Under torchrun with N ranks every rank renders its sample range
(64/N spp of every pixel, scene replicated) into its own {R,G,B,weight} film and the films are summed
onto rank 0 with ONE NCCL reduce of 132.7 MB per rank (iile_film_reduce_dev = Film::MergeFilmTile
across GPUs) INSIDE the timed region — total work is fixed: "strong" scaling.

  value     Msamples/s, whole job, scene resident in HBM, CUDA events on the launch stream around
            K steps, max over ranks; the weight channel of the reduced film is checked afterwards
            (sum == samples: no rank's film was lost)
  e2e       the same metric through the host-buffer call a user makes (iile_render at N = 1;
            iile_render_dev + iile_film_reduce_dev + read-back at N > 1), wall clock, with the film
            read back into page-locked host memory every step
  roofline  the traversal kernel (castKernel, 60-65 % of the step): algorithmic bytes of the rays the
            step traces = sum over rays of 32 B x nodes + 48 B x primitive tests + 48 B,
            the per-ray counts being the REFERENCE traversal's (instrumented oracle) on a sample of the
            render's own rays (iile_render_dump_rays_dev), divided by the kernel's device time in the
            timed steps (CUDA events around every launch, IILE_RENDER_TIME_KERNELS), against
            MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference (oracle/_ref: PathIntegrator::Render over its own BVHAccel) on
            this box's host cores, a bounded spp slice of the same image
  also      the other BASELINE configs, each with its own value / e2e / roofline / cpu_baseline:
              cfg2  configs[1]: 2^24 closest-hit + 2^24 any-hit casts per GPU on the killeroo BVH
                    (weak: every rank casts its own ray range, no collective)
              cfg3  configs[2]: room, 16 area lights, 1024x1024, maxdepth 8, 256 spp split over the ranks
              cfg1  configs[0]: killeroo-simple 400x400, maxdepth 5, 16 spp (and 256 spp) split likewise

`--impl reference` times the reference's own CPU implementation on the headline workload: the same
cfg5 scene, bounded to 1 spp of the full frame per step (per-spp cost is constant), all host cores,
rank 0 only.  The flat scene it reads is written by a separate process (bin/pbrt_b200 --flatten): the
reference arm's own process maps nothing of the product.
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
UNIT = "Msamples/s"
WORKLOADS = {
    # name: (spp of the whole image, description)
    "cfg5": (64, "cfg5: 8,000,000-triangle terrain, 3840x2160, path maxdepth 5, 64 spp (BASELINE configs[4])"),
    "cfg3": (256, "cfg3: room with 16 area lights, 1024x1024, path maxdepth 8, 256 spp (BASELINE configs[2])"),
    "cfg1": (16, "cfg1: killeroo-simple 400x400, path maxdepth 5, 16 spp (BASELINE configs[0])"),
}


def metric_name(workload):
    return f"Msamples/s path-traced, {workload}"


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
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled in-process
    (~1 ms per sample); the fields are the ones the profiling recipe's nvidia-smi line reports."""

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


def issue_view(t):
    """What the committed ncu capture says the kernel is really bound by: the share of issue slots used and of the 32
    lanes active in an issued instruction; their product is the fraction of the SM's lane-issue ceiling."""
    if not t or "issue" not in t:
        return None
    i = dict(t["issue"])
    i["frac_of_lane_issue_ceiling"] = i["issue_active_pct"] / 100.0 * i["active_lanes_per_instruction"] / 32.0
    return i


def measured_traffic(kernel):
    """profiles/traffic.json entry of a kernel: DRAM bytes (read + write) of the committed ncu capture,
    with what the capture covered (rays / launches), or None."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None
    return json.loads(p.read_text()).get(kernel)


# ---- scenes ------------------------------------------------------------------------------------
def terrain_scene_path(rank=0):
    """BASELINE configs[4] geometry: the seeded 8,000,000-triangle height field (tools/make_scenes.py
    cfg5), generated on this box (152 MB PLY, ~2 s)."""
    d = Path(f"/tmp/iile_cfg5_{os.getuid()}_{rank}")
    if not (d / "cfg5.pbrt").exists():
        subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_scenes.py"), "cfg5", str(d)],
                              stdout=subprocess.DEVNULL)
    return d / "cfg5.pbrt"


def scene_path(oracle, workload, rank=0):
    if workload in ("cfg5", "terrain"):
        return terrain_scene_path(rank)
    p = oracle.REF_SCENES / f"{workload}.pbrt"
    if not p.exists():
        raise SystemExit(f"bench.py: {p} is missing — run `make -C oracle ref` once where /root/reference "
                         "exists (the build products travel with the repo)")
    return p


def flatten_in_subprocess(pbrt_file, flat_file, spp=0):
    """The reference arm reads a flat scene (oracle/ref_harness rebuilds the reference's own objects from
    it).  It is written by the product's host front-end in a SEPARATE process, so that the process timing
    the reference maps nothing of the product."""
    exe = ROOT / "pbrt-v3-iile_b200" / "bin" / "pbrt_b200"
    cmd = [str(exe), "--flatten", str(flat_file)] + (["--spp", str(spp)] if spp else []) + [str(pbrt_file)]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)


# ---- the reference arm ----------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the headline workload, all host cores, bounded sample."""
    if rank != 0:
        return
    oracle = load_oracle()
    if not oracle.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return
    cores = os.cpu_count() or 1
    tmp = Path(f"/tmp/iile_bench_ref_{os.getuid()}")
    tmp.mkdir(exist_ok=True)
    wl = args.workload
    if wl == "cfg2":
        return run_reference_cfg2(args, oracle, tmp, world, cores)
    spp_total, desc = WORKLOADS[wl]
    spp_step = {"cfg5": 1, "cfg3": 4, "cfg1": 16}[wl]
    flat = tmp / f"{wl}.flat"
    flatten_in_subprocess(scene_path(oracle, wl), flat, spp_step)
    info = oracle.ref_render(flat, tmp / "ref.pfm", sampler="stock", repeats=args.warmup + args.steps)
    flat.unlink()
    secs = info["pass_seconds"][args.warmup:]
    t = float(np.mean(secs))
    value = info["samples"] / t / 1e6
    sample = (f"{spp_step} spp of the full {wl} frame per step (of {spp_total}; per-spp cost is constant), stock "
              f"RandomSampler, SamplerIntegrator::Render wall time, one BVH build for all steps")
    print(json.dumps({
        "impl": "reference", "metric": metric_name(wl), "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "description": desc, "spp_per_step": spp_step, "samples_per_step": info["samples"],
                   "same_image_as_b200_arm": True, "bounded_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                         "bvh_build_s": info["build_seconds"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference_cfg2(args, oracle, tmp, world, cores):
    flat_path = tmp / "cfg2.flat"
    flatten_in_subprocess(scene_path(oracle, "cfg2"), flat_path)
    flat = oracle.FlatScene(flat_path)
    lo, hi = oracle.killeroo_ray_box(flat)
    sample = 1 << 22
    rays = oracle.generate_rays(lo, hi, 0, sample)
    times = []
    for it in range(args.warmup + args.steps):
        _, a = oracle.ref_raycast(flat_path, rays, tmp, any_hit=False)
        _, b = oracle.ref_raycast(flat_path, rays, tmp, any_hit=True)
        if it >= args.warmup:
            times.append(a["seconds"] + b["seconds"])
    t = float(np.mean(times))
    value = 2 * sample / t / 1e6
    desc = "first 2^22 of the 2^24 cfg2 rays (same generator; a rate), closest-hit + any-hit, per step"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s BVH cast (closest-hit + any-hit), cfg2", "value": value,
        "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2", "rays_per_kind": sample, "bounded_sample": desc},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- B200 arm ----------------------------------------------------------------------------------------
class Ctx:
    """Per-process state of the B200 arm."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch, self.dist = torch, dist
        self.iile = load_product()
        if self.iile.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(local_rank)
        self.iile.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        self.oracle = load_oracle()      # scene paths + (rank 0) cpu_baseline / roofline counts only
        self.stream = torch.cuda.current_stream()
        self.st = self.stream.cuda_stream
        self.comm = self.iile.partition.film_comm(rank, world) if world > 1 else None
        self.tmp = Path(f"/tmp/iile_bench_{os.getuid()}_{rank}")
        self.tmp.mkdir(exist_ok=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        t = self.torch.tensor(xs, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]


def ray_bytes(oracle, flat, rays, kinds):
    """32 B x nodes fetched + 48 B x primitive tests by the REFERENCE traversal order
    (instrumented oracle) + 32 B ray + 16 B hit, averaged over the closest-hit rays (kinds 0-2) and over
    the any-hit rays (kind 3) of a sample of a render's own rays."""
    out = {}
    for name, sel, any_hit in (("closest", kinds < 3, False), ("shadow", kinds == 3, True)):
        r = rays[sel]
        if len(r) == 0:
            out[name] = {"bytes_per_ray": 48.0, "nodes": 0.0, "tests": 0.0, "rays": 0}
            continue
        _, nodes, tests = oracle.cast(flat, r, any_hit=any_hit)
        out[name] = {"bytes_per_ray": (32 * nodes + 48 * tests) / len(r) + 48, "nodes": nodes / len(r),
                     "tests": tests / len(r), "rays": int(len(r))}
    return out


def render_leg(cx: Ctx, workload, spp_total, steps, warm, *, cpu_spp=0, want_roofline=True):
    """One rendered BASELINE config: `spp_total` samples per pixel of the whole image split over the
    ranks by sample range, film reduced onto rank 0 with NCCL inside the timed region."""
    torch, iile, oracle = cx.torch, cx.iile, cx.oracle
    rank, world = cx.rank, cx.world
    path = scene_path(oracle, workload, 0) if (rank == 0 or world == 1) else None
    scene = iile.partition.load_scene_shared(iile.Scene, path, rank, world, upload=True)
    info = scene.info()
    base = info.setup
    setup = scene.info().setup
    setup.pixelSamples = spp_total
    scene.set_setup(setup)
    h, w = scene.film_shape()
    b, e = iile.partition.sample_range(spp_total, rank, world)
    film = torch.zeros(h, w, 4, device="cuda")
    timed = iile.RenderParams(b, e, 0, 0, 1, iile.RENDER_TIME_KERNELS, 0)
    plain = iile.RenderParams(b, e, 0, 0, 1, 0, 0)

    def step(params, want_stats):
        stats = None
        if e > b:
            stats = scene.render_dev(film.data_ptr(), params, cx.st, want_stats=want_stats)
        else:
            film.zero_()
        if cx.comm is not None:   # Film::MergeFilmTile across GPUs, on the render's stream
            cx.comm.reduce(film.data_ptr(), h * w, 0, cx.st)
        return stats

    for _ in range(max(warm, 1)):
        step(plain, False)
    # ---- device-timed: K steps between two events, barrier + synchronize on both sides ---------
    clocks = ClockSampler(cx.local_rank)
    cx.barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cx.stream)
    for _ in range(steps):
        step(plain, False)
    e1.record(cx.stream)
    cx.barrier()
    clk = clocks.stop()
    ms_step = cx.max_over_ranks(e0.elapsed_time(e1) / steps)
    # the reduced film holds every rank's samples: box-filter weights sum to the sample count
    samples_total = float(h) * w * spp_total
    weight_ok = True
    if rank == 0:
        wsum = float(film[..., 3].double().sum().item())
        # box filter: one pixel per sample, two when pFilm rounds onto a pixel edge (<= 1e-3 of the samples,
        # the reference's AddSample does the same); a lost rank would be 1/N of the weight
        weight_ok = 0 <= wsum / samples_total - 1 < 1e-3 if base.filterType == 0 else wsum > 0
        assert weight_ok, f"{workload}: film weight {wsum} != {samples_total} samples: a rank's film was lost"
        film_mean = float(scene.resolve(film.cpu().numpy()).mean())
    # ---- the same steps once more with events around every launch: kernel-family times ----------
    ksteps = max(1, min(steps, 2))
    fam = np.zeros(3)
    counts = np.zeros(4)
    for _ in range(ksteps):
        s = step(timed, True)
        if s is not None:
            fam += [s.castSeconds, s.shadeSeconds, s.otherSeconds]
            counts = np.array([s.cameraRays, s.closestRays, s.shadowRays, s.launches], dtype=np.float64)
    cx.barrier()
    fam /= ksteps
    cast_s = cx.max_over_ranks(fam[0])
    shade_s = cx.max_over_ranks(fam[1])
    other_s = cx.max_over_ranks(fam[2])
    camera, closest, shadow, launches = cx.sum_over_ranks(list(counts))
    # ---- end to end: the host-buffer call, film read back into page-locked memory every step -----
    pinned = torch.empty(h, w, 4, dtype=torch.float32, pin_memory=True)
    e2e_steps = max(1, min(steps, 3))

    def e2e_step():
        if world == 1:
            scene.render(plain, out=pinned.numpy())             # iile_render: kernels + D2H of the film
        else:
            step(plain, False)
            if rank == 0:
                pinned.copy_(film, non_blocking=True)
            torch.cuda.synchronize()

    e2e_step()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    cx.barrier()
    e2e_s = cx.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    out = None
    if rank == 0:
        value = samples_total / (ms_step * 1e-3) / 1e6
        out = {
            "metric": metric_name(workload), "value": value, "unit": UNIT, "ms_per_step": ms_step,
            "mrays_per_s": (closest + shadow) / (ms_step * 1e-3) / 1e6,
            "config": {"workload": workload, "description": WORKLOADS[workload][1],
                       "resolution": [int(setup.xRes), int(setup.yRes)], "spp_total": spp_total,
                       "spp_per_gpu": e - b, "maxdepth": int(setup.maxDepth), "triangles": int(info.nTriangles),
                       "nodes": int(info.nNodes), "lights": int(info.nLights),
                       "partition": "sample range per rank, scene replicated" +
                                    (f"; iile_film_reduce_dev (ncclReduce, NCCL {iile.FilmComm.nccl_version()}) of "
                                     f"{h * w * 16} bytes per rank inside the timed region" if world > 1 else ""),
                       "l2_policy": "working set larger than L2 between steps: path state + queues of 4 M paths "
                                    "(>= 800 MB) are rewritten by every batch",
                       "film_weight_sum_ok": weight_ok, "film_mean": film_mean, "bvh_build_s": float(info.buildSeconds)},
            "samples_per_step": samples_total, "rays_per_step": closest + shadow, "closest_rays": closest,
            "shadow_rays": shadow, "launches_per_step": launches / world, "gpu_launches": int(launches / world) * steps,
            "kernel_seconds_per_step": {"cast": cast_s, "shade": shade_s, "other": other_s,
                                        "note": "max over ranks; events around every launch in separate, untimed-for-value steps"},
            "e2e": {"value": samples_total / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 32,
                    "d2h_bytes_per_step": h * w * 16,
                    "note": "iile_render (host film)" if world == 1 else
                            "iile_render_dev + iile_film_reduce_dev + read-back of the merged film on rank 0; "
                            "input per step is the 32-byte IileRenderParams: the scene stays resident"},
            "clocks": clk,
        }
    # ---- roofline of the traversal kernel from a sample of the render's own rays (rank 0) ----------
    if rank == 0 and want_roofline:
        try:
            out["roofline"] = render_roofline(cx, scene, workload, out, cast_s, closest, shadow)
        except Exception as ex:   # the roofline is commentary: never lose the measurement over it
            out["roofline"] = {"error": repr(ex)}
    if rank == 0 and cpu_spp and oracle.have_ref() and not cx.args.no_cpu_baseline:
        s2 = scene.info().setup
        s2.pixelSamples = cpu_spp
        scene.set_setup(s2)
        flat = cx.tmp / "cpu.flat"
        scene.save_flat(flat)
        ref = oracle.ref_render(flat, cx.tmp / "ref.pfm", sampler="stock")
        flat.unlink()
        out["cpu_baseline"] = {"value": ref["msamples_per_s"], "unit": UNIT, "cores": ref["threads"], "kind": "reference",
                               "sample": f"{workload} at {cpu_spp} spp of the full frame (per-spp cost is constant), stock "
                                         "sampler, SamplerIntegrator::Render wall time, oracle/_ref",
                               "bvh_build_s": ref["build_seconds"]}
    scene.set_setup(base)
    cx.barrier()
    scene.close()
    return out


def render_roofline(cx, scene, workload, out, cast_s, closest, shadow):
    torch, iile, oracle = cx.torch, cx.iile, cx.oracle
    h, w = scene.film_shape()
    cap = 1 << 18
    setup = scene.info().setup
    spp = setup.pixelSamples
    setup.pixelSamples = 1
    scene.set_setup(setup)
    est = max(1.0, (closest + shadow) / max(out["samples_per_step"], 1.0) * h * w)   # rays of a 1-spp render
    stride = max(1, int(est / (cap * 0.8)))
    rays_d = torch.empty(cap * 8, dtype=torch.float32, device="cuda")
    kinds_d = torch.empty(cap, dtype=torch.uint8, device="cuda")
    scratch = torch.zeros(h, w, 4, device="cuda")
    offered = scene.render_dump_rays_dev(scratch.data_ptr(), iile.RenderParams(0, 1, 0, 0, 1, 0, 0),
                                         rays_d.data_ptr(), kinds_d.data_ptr(), cap, stride, cx.st)
    n = min(offered, cap)
    rays = rays_d.view(cap, 8)[:n].cpu().numpy().view(iile.RAY_DTYPE).reshape(-1)
    kinds = kinds_d[:n].cpu().numpy()
    setup.pixelSamples = spp
    scene.set_setup(setup)
    flat_path = cx.tmp / "roof.flat"
    scene.save_flat(flat_path)
    flat = oracle.FlatScene(flat_path)
    flat_path.unlink()
    rb = ray_bytes(oracle, flat, rays, kinds)
    total = closest * rb["closest"]["bytes_per_ray"] + shadow * rb["shadow"]["bytes_per_ray"]
    peak, src = measured_peak()
    achieved = total / cast_s / 1e9 / cx.world if cast_s > 0 else 0.0
    t = measured_traffic(f"castKernel@{workload}")
    traffic = None
    if t:   # DRAM bytes of the committed capture, scaled from the rays it covered to one step's rays per GPU
        traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["rays"] * (closest + shadow) / cx.world
    return {"bound": "hbm", "kernel": "castKernel (all launches of a step)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "issue": issue_view(t),
            "algorithmic_bytes_per_step_per_gpu": total / cx.world, "cast_seconds_per_step": cast_s,
            "share_of_step": cast_s / (out["ms_per_step"] * 1e-3), "peak_source": src,
            "per_ray": rb, "kind_mix_of_sample": {k: int((kinds == i).sum()) for i, k in
                                                  enumerate(("camera", "bounce", "mis", "shadow"))},
            "whole_step": {"achieved": (total + out["samples_per_step"] * 32) / (out["ms_per_step"] * 1e-3) / 1e9 / cx.world,
                           "frac": (total + out["samples_per_step"] * 32) / (out["ms_per_step"] * 1e-3) / 1e9 / cx.world / peak,
                           "note": "SURVEY 8d per-sample figure (rays + 32 B film) over the whole step, per GPU"},
            "note": "per-ray node / primitive-test counts: reference traversal order (instrumented oracle) on "
                    f"{n} rays sampled from this render (iile_render_dump_rays_dev, every {stride}th ray of a 1-spp pass); "
                    "queue and path-state traffic deliberately not credited"}


def cast_leg(cx: Ctx, workload="cfg2"):
    """BASELINE configs[1]: 2^24 closest-hit + 2^24 any-hit casts per GPU, rays sharded by range."""
    torch, iile, oracle = cx.torch, cx.iile, cx.oracle
    args, rank, world = cx.args, cx.rank, cx.world
    terrain = workload == "terrain"
    scene = iile.Scene.from_pbrt(scene_path(oracle, workload, rank), upload=True)
    info = scene.info()
    flat_path = cx.tmp / "cast.flat"
    scene.save_flat(flat_path)
    flat = oracle.FlatScene(flat_path)
    lo, hi = (oracle.cfg2_ray_box if terrain else oracle.killeroo_ray_box)(flat)
    n = N_RAYS
    steps, st = max(2, min(args.steps, 10)), cx.st
    rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
    hits = torch.empty(n * 4, dtype=torch.int32, device="cuda")
    occ = torch.empty(n, dtype=torch.uint8, device="cuda")
    iile.generate_rays_dev(lo, hi, rank * n, n, rays.data_ptr(), st)
    torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        scene.intersect_dev(rays.data_ptr(), n, hits.data_ptr(), st)
        scene.intersect_p_dev(rays.data_ptr(), n, occ.data_ptr(), st)
    clocks = ClockSampler(cx.local_rank)
    cx.barrier()
    clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cx.stream)
    for k in range(steps):
        ev[3 * k].record(cx.stream)
        scene.intersect_dev(rays.data_ptr(), n, hits.data_ptr(), st)
        ev[3 * k + 1].record(cx.stream)
        scene.intersect_p_dev(rays.data_ptr(), n, occ.data_ptr(), st)
        ev[3 * k + 2].record(cx.stream)
    e1.record(cx.stream)
    cx.barrier()
    clk = clocks.stop()
    ms_closest = float(np.mean([ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(steps)]))
    ms_any = float(np.mean([ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(steps)]))
    ms_step = cx.max_over_ranks(e0.elapsed_time(e1) / steps)
    value = world * 2 * n / (ms_step * 1e-3) / 1e6
    # ---- end to end: the host-buffer calls with pinned arrays; the rays are uploaded ONCE per cast kind
    h_rays = torch.empty(n * 8, dtype=torch.float32, pin_memory=True)
    h_rays.copy_(rays)
    h_hits = torch.empty(n * 4, dtype=torch.int32, pin_memory=True)
    h_occ = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    C = iile.C

    def e2e_step():
        iile._check(iile.lib.iile_intersect_closest(scene._h, C.c_void_p(h_rays.data_ptr()), n, C.c_void_p(h_hits.data_ptr())))
        iile._check(iile.lib.iile_intersect_any(scene._h, C.c_void_p(h_rays.data_ptr()), n, C.c_void_p(h_occ.data_ptr())))

    e2e_step()
    cx.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    cx.barrier()
    e2e_s = cx.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    nhit = int((hits.view(n, 4)[:, 0] >= 0).sum())
    assert int(occ.sum()) == nhit, "any-hit and closest-hit disagree"
    if rank != 0:
        scene.close()
        return None
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
    t = measured_traffic("closestHitKernel" if not terrain else "closestHitKernel@terrain")
    roofline = {"bound": "hbm", "kernel": "closestHitKernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None if t is None else float(t["dram_bytes_read"] + t["dram_bytes_write"]),
                "issue": issue_view(t), "algorithmic_bytes_per_launch": n * bytes_closest, "peak_source": peak_src,
                "bytes_per_ray": bytes_closest, "ms_per_launch": ms_closest,
                "any_hit": {"bytes_per_ray": bytes_any, "ms_per_launch": ms_any,
                            "achieved": n * bytes_any / (ms_any * 1e-3) / 1e9},
                "note": ("working set (3.8 MB nodes + 3.2 MB triangles) is L2-resident: instruction-issue bound, "
                         "not an HBM test (traffic << algorithmic bytes)") if not terrain else
                        ("working set ~%.0f MB of 4-wide nodes + %.0f MB of triangle slots: larger than the 126 MB L2"
                         % (info.nNodes / 4 * 128 / 1e6, info.nPrims * 48 / 1e6))}
    cpu = None
    if not args.no_cpu_baseline and oracle.have_ref():
        cs = 1 << 22
        r_c = rays.view(n, 8)[:cs].cpu().numpy().view(iile.RAY_DTYPE).reshape(-1)
        _, a = oracle.ref_raycast(flat_path, r_c, cx.tmp, any_hit=False)
        _, b = oracle.ref_raycast(flat_path, r_c, cx.tmp, any_hit=True)
        cpu = {"value": 2 * cs / (a["seconds"] + b["seconds"]) / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(),
               "kind": "reference", "sample": "first 2^22 of rank 0's rays, closest-hit + any-hit, "
               "oracle/_ref BVHAccel::Intersect/IntersectP via ParallelFor",
               "closest_mrays_s": a["mrays_per_s"], "any_mrays_s": b["mrays_per_s"]}
    scene.close()
    return {
        "metric": "Mrays/s BVH cast (closest-hit + any-hit), " + workload, "value": value, "unit": "Mrays/s",
        "ms_per_step": ms_step, "scaling": "weak", "steps": steps,
        "config": {"workload": workload, "rays_per_kind_per_gpu": n, "kinds": ["closest-hit", "any-hit"],
                   "triangles": int(info.nTriangles), "nodes": int(info.nNodes), "bvh": "SAH maxnodeprims 4",
                   "l2_policy": "inputs larger than L2 (512 MiB of rays per cast)",
                   "hit_fraction": nhit / n, "parity_vs_oracle_on_sample": parity},
        "e2e": {"value": world * 2 * n / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * n * 32,
                "d2h_bytes_per_step": n * 16 + n,
                "note": "iile_intersect_closest + iile_intersect_any, pinned host arrays; each call uploads its rays"},
        "gpu_launches": 2 * steps, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
    }


def run_b200(args, rank, world, local_rank):
    cx = Ctx(args, rank, world, local_rank)
    wl = args.workload
    steps, warm = args.steps, max(args.warmup, 3)
    if wl in ("cfg2", "terrain"):
        head = cast_leg(cx, wl)
        scaling = "weak"
    else:
        spp_total = args.spp or WORKLOADS[wl][0]
        head = render_leg(cx, wl, spp_total, steps, warm, cpu_spp={"cfg5": 2, "cfg3": 4, "cfg1": 16}[wl])
        scaling = "strong"
    also = {}
    for leg in [l for l in args.legs.split(",") if l and l != "none" and l != wl]:
        try:
            if leg in ("cfg2", "terrain"):
                also[leg] = cast_leg(cx, leg)
            elif leg == "cfg1":
                also["cfg1"] = render_leg(cx, "cfg1", 16, max(2, min(steps, 10)), 3, cpu_spp=16)
                also["cfg1_256spp"] = render_leg(cx, "cfg1", 256, max(2, min(steps, 3)), 1, want_roofline=False)
            else:
                also[leg] = render_leg(cx, leg, WORKLOADS[leg][0], max(2, min(steps, 3)), 1,
                                       cpu_spp={"cfg5": 2, "cfg3": 4}[leg])
        except Exception as ex:   # a side leg never takes the headline down
            if rank == 0:
                also[leg] = {"error": repr(ex)}
    if rank == 0:
        out = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        }
        for k in ("config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "mrays_per_s",
                  "samples_per_step", "rays_per_step", "closest_rays", "shadow_rays", "launches_per_step",
                  "kernel_seconds_per_step"):
            if k in head:
                out[k] = head[k]
        if also:
            out["also"] = also
        print(json.dumps(out))
    if world > 1:
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg3", "cfg1", "cfg2", "terrain"],
                    help="the headline: cfg5 = BASELINE configs[4] (default); cfg2 = the ray-cast microbenchmark; "
                         "terrain = cfg2's ray generator on the cfg5 geometry")
    ap.add_argument("--legs", default="cfg2,cfg3,cfg1", help="BASELINE configs reported under `also` (comma list, or none)")
    ap.add_argument("--spp", type=int, default=0, help="override the headline's samples per pixel (whole image)")
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
