"""oracle.py — TEST INFRASTRUCTURE (never imported by the product package).

Python face of the two checkers:
  * liboracle.so  — our scalar C++ restatement (oracle/cast_oracle.cpp)
  * oracle/_ref/  — the unmodified reference, compiled by oracle/Makefile, driven through
                    oracle/ref_harness (BVHAccel / PathIntegrator on a flat scene) or as the
                    stock `pbrt` binary.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
REF_HARNESS = REF_DIR / "ref_harness"
REF_PBRT = REF_DIR / "pbrt"
REF_SCENES = REF_DIR / "scenes"
LIB = HERE / "_build" / "liboracle.so"

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tMax", "<f4"), ("d", "<f4", 3), ("time", "<f4")])
HIT_DTYPE = np.dtype([("prim", "<i4"), ("t", "<f4"), ("b1", "<f4"), ("b2", "<f4")])
REFHIT_DTYPE = np.dtype([("prim", "<i4"), ("t", "<f4"), ("b1", "<f4"), ("b2", "<f4"),
                         ("p", "<f4", 3), ("n", "<f4", 3), ("pad", "<f4", 2)])
NODE_DTYPE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("offset", "<i4"),
                       ("nPrims", "<u2"), ("axis", "u1"), ("pad", "u1")])
PRIM_DTYPE = np.dtype([("type", "<i4"), ("v", "<i4", 3), ("material", "<i4"), ("light", "<i4"),
                       ("flags", "<u4"), ("pad", "<u4")])
SPHERE_DTYPE = np.dtype([("o2w", "<f4", 16), ("w2o", "<f4", 16), ("radius", "<f4"), ("zMin", "<f4"),
                         ("zMax", "<f4"), ("thetaMin", "<f4"), ("thetaMax", "<f4"), ("phiMax", "<f4"),
                         ("flags", "<u4"), ("phiMaxDeg", "<f4")])
MATERIAL_BYTES = 96      # IILEFLT2 (ABI 2); IILEFLT1 files carry 64-byte materials
LIGHT_BYTES = 272        # IILEFLT2; IILEFLT1 files carry 40-byte lights
SETUP_BYTES = 520


def build(force=False):
    """Compile the restatement (and, when /root/reference is present, oracle/_ref)."""
    if force or not LIB.exists() or LIB.stat().st_mtime < (HERE / "cast_oracle.cpp").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "restate"], stdout=subprocess.DEVNULL)
    if Path("/root/reference/src").exists():
        subprocess.check_call(["make", "-C", str(HERE), "-j8", "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        _lib = C.CDLL(str(LIB))
    return _lib


class FlatScene:
    """Reader for the IILEFLT1 / IILEFLT2 container (csrc/host/flat_io.cpp)."""

    def __init__(self, path):
        raw = Path(path).read_bytes()
        assert raw[:8] in (b"IILEFLT1", b"IILEFLT2"), "not a flat scene"
        mat_bytes = 64 if raw[:8] == b"IILEFLT1" else MATERIAL_BYTES
        c = np.frombuffer(raw, dtype="<u4", count=8, offset=8)
        nv, hasN, hasUV, nPrims, nSph, nMat, nLights, nNodes = [int(x) for x in c]
        off = 40
        self.setup_raw = raw[off:off + SETUP_BYTES]
        off += SETUP_BYTES

        def take(dtype, count):
            nonlocal off
            a = np.frombuffer(raw, dtype=dtype, count=count, offset=off).copy()
            off += a.nbytes
            return a

        self.P = take("<f4", nv * 3).reshape(-1, 3)
        self.N = take("<f4", nv * 3 if hasN else 0).reshape(-1, 3)
        self.UV = take("<f4", nv * 2 if hasUV else 0).reshape(-1, 2)
        self.prims = take(PRIM_DTYPE, nPrims)
        self.spheres = take(SPHERE_DTYPE, nSph)
        self.materials = take(np.dtype(("u1", mat_bytes)), nMat)
        self.lights = take(np.dtype(("u1", 40 if raw[:8] == b"IILEFLT1" else LIGHT_BYTES)), nLights)
        self.nodes = take(NODE_DTYPE, nNodes)
        self.prim_order = take("<u4", nPrims if nNodes else 0)
        assert off == len(raw), (off, len(raw))

    def triangle_bounds(self):
        tri = self.prims[self.prims["type"] == 0]
        v = self.P[tri["v"].reshape(-1)]
        return v.min(axis=0), v.max(axis=0)


def cfg2_ray_box(flat: FlatScene):
    """Origin box of the cfg-2 ray set: triangle bound grown by half its diagonal."""
    lo, hi = flat.triangle_bounds()
    half = (hi - lo) * np.float32(0.5)
    return (lo - half).astype(np.float32), (hi + half).astype(np.float32)


def killeroo_ray_box(flat: FlatScene):
    """bound of the killeroo triangles alone (the loop-subdivision meshes, which
    are the ones carrying normals), grown by half its diagonal on every side."""
    tri = flat.prims[(flat.prims["type"] == 0) & ((flat.prims["flags"] & 1) != 0)]
    v = flat.P[tri["v"].reshape(-1)]
    lo, hi = v.min(axis=0), v.max(axis=0)
    half = (hi - lo) * np.float32(0.5)
    return (lo - half).astype(np.float32), (hi + half).astype(np.float32)


def generate_rays(lo, hi, first, n) -> np.ndarray:
    rays = np.empty(n, dtype=RAY_DTYPE)
    lo = np.ascontiguousarray(lo, dtype=np.float32)
    hi = np.ascontiguousarray(hi, dtype=np.float32)
    lib().oracle_generate_rays(lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                               C.c_uint64(first), C.c_uint64(n), rays.ctypes.data_as(C.c_void_p))
    return rays


def pcg32_floats(seq, skip, n) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    lib().oracle_pcg32_floats(C.c_uint64(seq), C.c_uint64(skip), C.c_uint64(n), out.ctypes.data_as(C.c_void_p))
    return out


def cast(flat: FlatScene, rays: np.ndarray, any_hit=False, threads=0):
    """Restated BVHAccel::Intersect / IntersectP.  Returns (hits|occluded, nodes_fetched, prim_tests)."""
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    n = rays.shape[0]
    hits = np.zeros(0 if any_hit else n, dtype=HIT_DTYPE)
    occ = np.zeros(n if any_hit else 0, dtype=np.uint8)
    counts = np.zeros(2, dtype=np.uint64)
    nodes = np.ascontiguousarray(flat.nodes)
    P = np.ascontiguousarray(flat.P)
    slot = np.ascontiguousarray(flat.prim_order.astype(np.int32))
    ptype = np.ascontiguousarray(flat.prims["type"].astype(np.int32))
    pv = np.ascontiguousarray(flat.prims["v"].astype(np.int32))
    sph = np.ascontiguousarray(flat.spheres)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().oracle_cast(vp(nodes), vp(P), vp(slot), vp(ptype), vp(pv), vp(sph), vp(rays), C.c_uint64(n),
                      C.c_int(1 if any_hit else 0), vp(hits), vp(occ), vp(counts), C.c_int(threads))
    return (occ if any_hit else hits), int(counts[0]), int(counts[1])


# ---- the reference itself -------------------------------------------------------------
def have_ref() -> bool:
    return REF_HARNESS.exists() and REF_PBRT.exists()


def _run(cmd, timeout=3600):
    p = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"{cmd[0]} failed ({p.returncode}): {p.stderr[-2000:]}")
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    return json.loads(line[-1]) if line else {}


def ref_bvh(flat_path, workdir):
    wd = Path(workdir)
    info = _run([REF_HARNESS, "bvh", flat_path, wd / "ref_nodes.bin", wd / "ref_order.bin"])
    return (np.fromfile(wd / "ref_nodes.bin", dtype=NODE_DTYPE),
            np.fromfile(wd / "ref_order.bin", dtype="<u4"), info)


def ref_raycast(flat_path, rays: np.ndarray, workdir, any_hit=False, threads=0):
    wd = Path(workdir)
    np.ascontiguousarray(rays, dtype=RAY_DTYPE).tofile(wd / "rays.bin")
    info = _run([REF_HARNESS, "raycast", flat_path, wd / "rays.bin", wd / "hits.bin",
                 "any" if any_hit else "closest", threads])
    out = np.fromfile(wd / "hits.bin", dtype=np.uint8 if any_hit else REFHIT_DTYPE)
    return out, info


def ref_render(flat_path, out_pfm, sampler="pixelstream", threads=0, spp=0, sample_offset=0, timeout=7200, repeats=1):
    """`repeats` > 1: that many Render() passes over one built scene; info["pass_seconds"] lists them."""
    return _run([REF_HARNESS, "render", flat_path, out_pfm, sampler, threads, spp, sample_offset, repeats], timeout)


def ref_hemispheres(flat_path, pos, direction, workdir, size=32, spp=1, threads=0):
    """IISPTdIntegrator::RenderView of the reference for a batch of hemispheric cameras.
    Returns (intensity[n,size,size,3], normals[n,size,size,3], distance[n,size,size], info);
    row y / column x = camera pixel (x, y)."""
    wd = Path(workdir)
    pts = np.concatenate([np.asarray(pos, np.float32).reshape(-1, 3),
                          np.asarray(direction, np.float32).reshape(-1, 3)], axis=1)
    pts.astype("<f4").tofile(wd / "hemi_points.bin")
    info = _run([REF_HARNESS, "hemi", flat_path, wd / "hemi_points.bin", wd / "hemi", size, spp, threads])
    n = len(pts)
    inten = np.stack([read_pfm(wd / f"hemi_d_{h}.pfm") for h in range(n)])
    nrm = np.stack([read_pfm(wd / f"hemi_n_{h}.pfm") for h in range(n)])
    dist = np.stack([read_pfm(wd / f"hemi_z_{h}.pfm")[..., 0] for h in range(n)])
    return inten, nrm, dist, info


def ref_hemispheres_nn(flat_path, pos, direction, workdir, size=32, spp=1, threads=0):
    """The reference's IisptRenderRunner::normalizeMapsDownstream / transformMapsUpstream
    (iisptrenderrunner.cpp:1041-1136) around RenderView, for a batch of hemispheric cameras
    (ref_harness hemi_nn).  Returns (nn_input[n,7,size,size] — the tensor main_stdio_net.read_input
    builds from what IisptNnConnector pipes: planes R,G,B, nx,ny,nz, distance, ImageFilm row order —,
    means[n,3], upstream[n,size,size,3] = transformMapsUpstream of the normalised intensity)."""
    wd = Path(workdir)
    pts = np.concatenate([np.asarray(pos, np.float32).reshape(-1, 3),
                          np.asarray(direction, np.float32).reshape(-1, 3)], axis=1)
    pts.astype("<f4").tofile(wd / "hemi_points.bin")
    _run([REF_HARNESS, "hemi_nn", flat_path, wd / "hemi_points.bin", wd / "heminn", size, spp, threads])
    nn, means, up = [], [], []
    for h in range(len(pts)):
        a = np.fromfile(wd / f"heminn_nn_{h}.bin", dtype="<f4")
        px = size * size
        inten = a[:3 * px].reshape(size, size, 3).transpose(2, 0, 1)
        nrm = a[3 * px:6 * px].reshape(size, size, 3).transpose(2, 0, 1)
        dist = a[6 * px:7 * px].reshape(1, size, size)
        nn.append(np.concatenate([inten, nrm, dist], axis=0))
        means.append(a[7 * px:7 * px + 3])
        up.append(np.fromfile(wd / f"heminn_up_{h}.bin", dtype="<f4").reshape(size, size, 3))
    return np.stack(nn), np.stack(means), np.stack(up)


def ref_direct_pass(flat_path, workdir, passes):
    """IILE's direct-illumination pass: the reference's DirectProgressiveIntegrator::RenderOnePass `passes` times into
    its IisptFilmMonitor (ref_harness direct_pass).  Returns (rgb [h, w, 3] over Film::GetSampleBounds(), info)."""
    wd = Path(workdir)
    out = wd / "direct_pass.bin"
    info = _run([REF_HARNESS, "direct_pass", flat_path, out, passes])
    w, h = np.fromfile(out, dtype="<i4", count=2)
    return np.fromfile(out, dtype="<f4", offset=8).reshape(h, w, 3), info


def ref_iile_render(flat_path, workdir, indirect_tasks, direct_samples, size, radius_start, seed, timeout=3600):
    """IISPTIntegrator::render_normal_2 from the reference's own parts on one thread (ref_harness iile_render): the
    schedule monitor's tasks with the identity network, the direct passes, merge_into.  Returns (rgb [h, w, 3], info)."""
    wd = Path(workdir)
    out = wd / "iile_render.bin"
    info = _run([REF_HARNESS, "iile_render", flat_path, out, indirect_tasks, direct_samples, size, radius_start, seed], timeout)
    w, h = np.fromfile(out, dtype="<i4", count=2)
    return np.fromfile(out, dtype="<f4", offset=8).reshape(h, w, 3), info


def ref_iispt_task(flat_path, workdir, x0, y0, x1, y1, tilesize, size=32, sampler_counter=0, sampler_sample=0,
                   rng_seed=0, films=None, spp=1, timeout=3600):
    """One task of IisptRenderRunner::run walked by the reference's own members (ref_harness iispt_task).
    films: None = the identity network on the reference's own RenderView, else float32 [nValid,size,size,3]
    handed to HemisphericCamera::set_nn_film.  Returns (hemi [n,7] = valid, pos, dir; films
    [nValid,size,size,3]; pixels [y1-y0, x1-x0, 4] = f_beta*L and the sample weight; info)."""
    wd = Path(workdir)
    films_arg = "-"
    if films is not None:
        films_arg = wd / "iispt_given_films.bin"
        np.ascontiguousarray(films, dtype="<f4").tofile(films_arg)
    info = _run([REF_HARNESS, "iispt_task", flat_path, wd / "iispt", x0, y0, x1, y1, tilesize, size, sampler_counter,
                 sampler_sample, rng_seed, films_arg, spp], timeout)
    hemi = np.fromfile(wd / "iispt_hemi.bin", dtype="<f4").reshape(-1, 7)
    out_films = np.fromfile(wd / "iispt_films.bin", dtype="<f4").reshape(-1, size, size, 3)
    pixels = np.fromfile(wd / "iispt_pixels.bin", dtype="<f4").reshape(y1 - y0, x1 - x0, 4)
    return hemi, out_films, pixels, info


def ref_stats(out_pfm) -> dict:
    """The reference's own counters (stats.h) written by `ref_harness render` next to the image."""
    import re
    txt = Path(str(out_pfm) + ".stats.txt").read_text()
    out = {}
    for key, pat in (("cameraRays", r"Camera rays traced\s+(\d+)"),
                     ("closestRays", r"Regular ray intersection tests\s+(\d+)"),
                     ("shadowRays", r"Shadow ray intersection tests\s+(\d+)")):
        m = re.search(pat, txt)
        out[key] = int(m.group(1)) if m else 0      # stats.h prints no line for a counter that stayed at zero
    return out


def read_pfm(path) -> np.ndarray:
    with open(path, "rb") as f:
        kind = f.readline().strip()
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        ch = 3 if kind == b"PF" else 1
        a = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, ch)
    return a[::-1].astype(np.float32)
