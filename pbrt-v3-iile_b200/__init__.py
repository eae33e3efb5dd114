"""Python host mirror of the B200 rendering hot path of pbrt-v3-IILE.

Thin ctypes binding over the C ABI in ``include/iile_b200.h``
(``lib/libiile_b200.so``, built by ``csrc/Makefile``).  Names follow the
reference's API for this path: ``Scene.from_pbrt`` ~ ParseFile + MakeScene
(src/core/api.cpp:1632-1703), ``Scene.intersect`` / ``intersect_p`` ~
BVHAccel::Intersect / IntersectP (src/accelerators/bvh.cpp:662-738),
``Scene.render`` ~ SamplerIntegrator::Render (src/core/integrator.cpp:227-339).

There is no CPU fallback: if the shared library is missing, importing raises;
if no CUDA device is present every compute call raises ``IileError``.
PyTorch is used by callers only for device buffers / streams / NCCL; this
module needs numpy alone and accepts raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("IILE_B200_LIB", _HERE / "lib" / "libiile_b200.so"))  # override: kernel tuning builds


E_UNSUPPORTED = -5  # IILE_E_UNSUPPORTED, include/iile_b200.h


class IileError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[iile {code}] {msg}")
        self.code = code


if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C {_HERE / 'csrc'}` "
        "(or __graft_entry__.build()); there is no fallback path")

lib = C.CDLL(str(LIB_PATH))


# ---- POD mirrors of include/iile_b200.h --------------------------------------
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tMax", "<f4"), ("d", "<f4", 3), ("time", "<f4")])
HIT_DTYPE = np.dtype([("prim", "<i4"), ("t", "<f4"), ("b1", "<f4"), ("b2", "<f4")])
NODE_DTYPE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("offset", "<i4"),
                       ("nPrims", "<u2"), ("axis", "u1"), ("pad", "u1")])
PRIM_DTYPE = np.dtype([("type", "<i4"), ("v", "<i4", 3), ("material", "<i4"), ("light", "<i4"),
                       ("flags", "<u4"), ("pad", "<u4")])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16 and NODE_DTYPE.itemsize == 32
assert PRIM_DTYPE.itemsize == 32


class RenderSetup(C.Structure):
    _fields_ = [
        ("cameraToWorld", C.c_float * 16), ("worldToCamera", C.c_float * 16),
        ("fov", C.c_float), ("lensRadius", C.c_float), ("focalDistance", C.c_float),
        ("screenWindow", C.c_float * 4), ("shutterOpen", C.c_float), ("shutterClose", C.c_float),
        ("xRes", C.c_int32), ("yRes", C.c_int32), ("cropWindow", C.c_float * 4),
        ("filmScale", C.c_float), ("maxSampleLuminance", C.c_float), ("diagonal", C.c_float),
        ("filterType", C.c_int32), ("filterRadius", C.c_float * 2), ("filterParam", C.c_float * 2),
        ("pixelSamples", C.c_int32), ("maxDepth", C.c_int32), ("rrThreshold", C.c_float),
        ("hasPixelBounds", C.c_int32), ("pixelBounds", C.c_int32 * 4),
        ("lightSampleStrategy", C.c_int32), ("maxNodePrims", C.c_int32), ("splitMethod", C.c_int32),
        ("filename", C.c_char * 236), ("cameraType", C.c_int32), ("integratorType", C.c_int32), ("directStrategy", C.c_int32),
        ("samplerType", C.c_int32), ("samplePixelCenter", C.c_int32),
    ]


SAMPLER_RANDOM, SAMPLER_HALTON, SAMPLER_SOBOL = 0, 1, 2
FILTER_BOX, FILTER_GAUSSIAN, FILTER_MITCHELL, FILTER_SINC, FILTER_TRIANGLE, FILTER_MONITOR = range(6)


class SceneInfo(C.Structure):
    _fields_ = [
        ("nVerts", C.c_uint32), ("nPrims", C.c_uint32), ("nTriangles", C.c_uint32),
        ("nSpheres", C.c_uint32), ("nMaterials", C.c_uint32), ("nLights", C.c_uint32),
        ("nNodes", C.c_uint32), ("worldMin", C.c_float * 3), ("worldMax", C.c_float * 3),
        ("setup", RenderSetup), ("sampleBounds", C.c_int32 * 4), ("croppedBounds", C.c_int32 * 4),
        ("buildSeconds", C.c_double),
    ]


class SceneDesc(C.Structure):
    """IileSceneDesc: caller-owned arrays of a flattened scene (include/iile_b200.h)."""
    _fields_ = [
        ("nVerts", C.c_uint32), ("P", C.c_void_p), ("N", C.c_void_p), ("UV", C.c_void_p),
        ("nPrims", C.c_uint32), ("prims", C.c_void_p),
        ("nSpheres", C.c_uint32), ("spheres", C.c_void_p),
        ("nMaterials", C.c_uint32), ("materials", C.c_void_p),
        ("nLights", C.c_uint32), ("lights", C.c_void_p),
        ("nNodes", C.c_uint32), ("nodes", C.c_void_p), ("primOrder", C.c_void_p),
        ("setup", RenderSetup),
        # ABI 4: textures (all optional)
        ("nTextures", C.c_uint32), ("textures", C.c_void_p),
        ("nImages", C.c_uint32), ("images", C.c_void_p),
        ("nTexels", C.c_uint64), ("texels", C.c_void_p),
        ("nMaterialTextures", C.c_uint32), ("materialTextures", C.c_void_p),
    ]


MATERIAL_BYTES, LIGHT_BYTES, SPHERE_BYTES = 96, 272, 160
TEXTURE_BYTES, IMAGE_BYTES, MATERIAL_TEXTURES_BYTES = 64, 32, 48


class RenderParams(C.Structure):
    _fields_ = [("sampleBegin", C.c_int32), ("sampleEnd", C.c_int32), ("tileBegin", C.c_int32),
                ("tileEnd", C.c_int32), ("clearFilm", C.c_int32), ("flags", C.c_int32),
                ("maxPathsInFlight", C.c_uint64)]


class RenderStats(C.Structure):
    _fields_ = [("cameraRays", C.c_uint64), ("closestRays", C.c_uint64), ("shadowRays", C.c_uint64),
                ("launches", C.c_uint64), ("deviceSeconds", C.c_double),
                ("castSeconds", C.c_double), ("shadeSeconds", C.c_double), ("otherSeconds", C.c_double)]


RENDER_TIME_KERNELS = 1


class IisptTask(C.Structure):
    """IileIisptTask: one IisptScheduleMonitorTask of IisptRenderRunner::run (iisptrenderrunner.cpp:216-596)."""
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("tilesize", C.c_int32), ("hemiSize", C.c_int32), ("samplerCounter", C.c_uint32),
                ("samplerSample", C.c_uint32), ("rngSeed", C.c_uint64)]


_vp = C.c_void_p
lib.iile_last_error.restype = C.c_char_p
lib.iile_abi_version.restype = C.c_int
lib.iile_device_count.restype = C.c_int
for _name, _args in {
    "iile_set_device": [C.c_int],
    "iile_scene_load_pbrt": [C.c_char_p, C.c_int, C.POINTER(_vp)],
    "iile_scene_load_flat": [C.c_char_p, C.c_int, C.POINTER(_vp)],
    "iile_scene_create": [C.POINTER(SceneDesc), C.c_int, C.POINTER(_vp)],
    "iile_scene_save_flat": [_vp, C.c_char_p],
    "iile_scene_info": [_vp, C.POINTER(SceneInfo)],
    "iile_scene_get_nodes": [_vp, _vp, C.c_uint32],
    "iile_scene_get_prim_order": [_vp, _vp, C.c_uint32],
    "iile_scene_set_setup": [_vp, C.POINTER(RenderSetup)],
    "iile_intersect_closest": [_vp, _vp, C.c_size_t, _vp],
    "iile_intersect_any": [_vp, _vp, C.c_size_t, _vp],
    "iile_intersect_closest_dev": [_vp, _vp, C.c_size_t, _vp, _vp],
    "iile_intersect_any_dev": [_vp, _vp, C.c_size_t, _vp, _vp],
    "iile_generate_rays_dev": [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint64, C.c_size_t, _vp, _vp],
    "iile_render_dev": [_vp, C.POINTER(RenderParams), _vp, C.POINTER(RenderStats), _vp],
    "iile_render": [_vp, C.POINTER(RenderParams), _vp, C.POINTER(RenderStats)],
    "iile_render_dump_rays_dev": [_vp, C.POINTER(RenderParams), _vp, _vp, _vp, C.c_uint64, C.c_uint32,
                                  C.POINTER(C.c_uint64), _vp],
    "iile_film_resolve": [_vp, _vp, _vp],
    "iile_write_image": [C.c_char_p, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    "iile_render_hemispheres": [_vp, C.c_size_t, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp,
                                C.POINTER(RenderStats)],
    "iile_hemispheres_to_nn_dev": [_vp, C.c_size_t, _vp, _vp, C.c_int, C.c_int, _vp, _vp,
                                   C.POINTER(RenderStats), _vp],
    "iile_hemispheres_from_nn_dev": [C.c_size_t, C.c_int, _vp, _vp, _vp, _vp],
    "iile_iispt_grid": [C.POINTER(IisptTask), C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "iile_iispt_hemi_points": [_vp, C.POINTER(IisptTask), _vp, _vp, _vp],
    "iile_iispt_gather_dev": [_vp, C.POINTER(IisptTask), _vp, _vp, _vp, _vp, _vp, _vp],
    "iile_dev_malloc": [C.c_size_t, C.POINTER(_vp)],
    "iile_dev_free": [_vp],
    "iile_dev_upload": [_vp, _vp, C.c_size_t],
    "iile_dev_download": [_vp, _vp, C.c_size_t],
    "iile_dev_synchronize": [_vp],
    "iile_comm_nccl_version": [C.POINTER(C.c_int)],
    "iile_comm_unique_id": [_vp],
    "iile_comm_create": [_vp, C.c_int, C.c_int, C.POINTER(_vp)],
    "iile_comm_destroy": [_vp],
    "iile_film_reduce_dev": [_vp, C.c_size_t, C.c_int, _vp, _vp],
}.items():
    fn = getattr(lib, _name)
    fn.argtypes = _args
    fn.restype = C.c_int
lib.iile_scene_destroy.argtypes = [_vp]
lib.iile_scene_destroy.restype = None

EXPORTED_SYMBOLS = [
    "iile_last_error", "iile_abi_version", "iile_device_count", "iile_set_device",
    "iile_scene_load_pbrt", "iile_scene_create", "iile_scene_save_flat", "iile_scene_load_flat",
    "iile_scene_destroy", "iile_scene_info", "iile_scene_get_nodes", "iile_scene_get_prim_order",
    "iile_scene_set_setup", "iile_intersect_closest", "iile_intersect_any",
    "iile_intersect_closest_dev", "iile_intersect_any_dev", "iile_generate_rays_dev",
    "iile_render_dev", "iile_render", "iile_render_dump_rays_dev", "iile_film_resolve", "iile_write_image",
    "iile_render_hemispheres", "iile_hemispheres_to_nn_dev", "iile_hemispheres_from_nn_dev",
    "iile_iispt_grid", "iile_iispt_hemi_points", "iile_iispt_gather_dev",
    "iile_dev_malloc", "iile_dev_free", "iile_dev_upload", "iile_dev_download", "iile_dev_synchronize",
    "iile_comm_nccl_version", "iile_comm_unique_id", "iile_comm_create", "iile_comm_destroy",
    "iile_film_reduce_dev",
]


def _check(rc: int):
    if rc != 0:
        raise IileError(rc, (lib.iile_last_error() or b"").decode(errors="replace"))


def device_count() -> int:
    return lib.iile_device_count()


def set_device(i: int):
    _check(lib.iile_set_device(i))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


class Scene:
    """A flattened scene + its SAH BVH, optionally resident on the current GPU."""

    def __init__(self, handle):
        self._h = handle

    # -- construction --------------------------------------------------------
    @classmethod
    def from_pbrt(cls, path, upload=True):
        h = _vp()
        _check(lib.iile_scene_load_pbrt(os.fsencode(str(path)), int(upload), C.byref(h)))
        return cls(h)

    @classmethod
    def from_flat(cls, path, upload=True):
        h = _vp()
        _check(lib.iile_scene_load_flat(os.fsencode(str(path)), int(upload), C.byref(h)))
        return cls(h)

    @classmethod
    def from_desc(cls, P, prims, setup, *, N=None, UV=None, spheres=None, materials=None, lights=None,
                  nodes=None, prim_order=None, upload=True):
        """iile_scene_create: the scene as the arrays a host that keeps its own parser would hand over
        (INTEGRATION.md).  P [nVerts,3] float32 world space; prims PRIM_DTYPE; spheres / materials /
        lights as raw records (uint8 [n, SPHERE_BYTES / MATERIAL_BYTES / LIGHT_BYTES]); optionally the
        caller's own LinearBVHNode array + primitive order, otherwise the BVH is built here."""
        keep = []

        def arr(a, dtype, width=None):
            if a is None:
                return None, 0
            a = np.ascontiguousarray(a, dtype=dtype)
            if width is not None and (a.ndim != 2 or a.shape[1] != width):
                raise ValueError("expected records of %d bytes" % width)
            keep.append(a)
            return _ptr(a), len(a)

        d = SceneDesc()
        d.P, d.nVerts = arr(np.asarray(P, np.float32).reshape(-1, 3), np.float32)
        d.N, _ = arr(None if N is None or len(N) == 0 else np.asarray(N, np.float32).reshape(-1, 3), np.float32)
        d.UV, _ = arr(None if UV is None or len(UV) == 0 else np.asarray(UV, np.float32).reshape(-1, 2), np.float32)
        d.prims, d.nPrims = arr(prims, PRIM_DTYPE)
        d.spheres, d.nSpheres = arr(spheres, np.uint8, SPHERE_BYTES)
        d.materials, d.nMaterials = arr(materials, np.uint8, MATERIAL_BYTES)
        d.lights, d.nLights = arr(lights, np.uint8, LIGHT_BYTES)
        d.nodes, d.nNodes = arr(nodes, NODE_DTYPE)
        d.primOrder, _ = arr(prim_order, np.uint32)
        d.setup = setup
        h = _vp()
        _check(lib.iile_scene_create(C.byref(d), int(upload), C.byref(h)))
        return cls(h)

    def save_flat(self, path):
        _check(lib.iile_scene_save_flat(self._h, os.fsencode(str(path))))

    def close(self):
        if self._h:
            lib.iile_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inspection ----------------------------------------------------------
    def info(self) -> SceneInfo:
        out = SceneInfo()
        _check(lib.iile_scene_info(self._h, C.byref(out)))
        return out

    def nodes(self) -> np.ndarray:
        n = self.info().nNodes
        a = np.zeros(n, dtype=NODE_DTYPE)
        _check(lib.iile_scene_get_nodes(self._h, _ptr(a), n))
        return a

    def prim_order(self) -> np.ndarray:
        n = self.info().nPrims
        a = np.zeros(n, dtype=np.uint32)
        _check(lib.iile_scene_get_prim_order(self._h, _ptr(a), n))
        return a

    def set_setup(self, setup: RenderSetup):
        _check(lib.iile_scene_set_setup(self._h, C.byref(setup)))

    # -- BVHAccel::Intersect / IntersectP -------------------------------------
    def intersect(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        _check(lib.iile_intersect_closest(self._h, _ptr(rays), rays.shape[0], _ptr(hits)))
        return hits

    def intersect_p(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        _check(lib.iile_intersect_any(self._h, _ptr(rays), rays.shape[0], _ptr(occ)))
        return occ

    def intersect_dev(self, rays_ptr: int, n: int, hits_ptr: int, stream: int = 0):
        _check(lib.iile_intersect_closest_dev(self._h, _vp(rays_ptr), n, _vp(hits_ptr), _vp(stream)))

    def intersect_p_dev(self, rays_ptr: int, n: int, occ_ptr: int, stream: int = 0):
        _check(lib.iile_intersect_any_dev(self._h, _vp(rays_ptr), n, _vp(occ_ptr), _vp(stream)))

    # -- SamplerIntegrator::Render ---------------------------------------------
    def film_shape(self):
        i = self.info()
        cb = i.croppedBounds
        return (cb[3] - cb[1], cb[2] - cb[0])

    def render(self, params: RenderParams | None = None, out: np.ndarray | None = None):
        """Render into a host film; returns (filmRGBW[h,w,4], RenderStats).  `out` reuses a caller-owned
        float32 [h,w,4] buffer (page-locked memory is read back by direct DMA)."""
        h, w = self.film_shape()
        if out is None:
            film = np.zeros((h, w, 4), dtype=np.float32)
        else:
            film = out
            if film.shape != (h, w, 4) or film.dtype != np.float32 or not film.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous float32 array of shape (%d, %d, 4)" % (h, w))
        stats = RenderStats()
        p = params or RenderParams(0, 0, 0, 0, 1, 0, 0)
        _check(lib.iile_render(self._h, C.byref(p), _ptr(film), C.byref(stats)))
        return film, stats

    def render_dev(self, film_ptr: int, params: RenderParams, stream: int = 0, want_stats: bool = True):
        """Enqueue a render into a device film on `stream`.  With want_stats the call synchronises
        the stream once at the end to read the ray counters back; without it stays asynchronous."""
        stats = RenderStats() if want_stats else None
        _check(lib.iile_render_dev(self._h, C.byref(params), _vp(film_ptr),
                                   C.byref(stats) if want_stats else None, _vp(stream)))
        return stats

    def render_dump_rays_dev(self, film_ptr: int, params: RenderParams, rays_ptr: int, kinds_ptr: int, cap: int,
                             stride: int = 1, stream: int = 0) -> int:
        """iile_render_dump_rays_dev (measurement aid): renders and copies every `stride`-th traced ray
        into the device buffers; returns the number of rays offered (min(that, cap) were stored)."""
        count = C.c_uint64(0)
        _check(lib.iile_render_dump_rays_dev(self._h, C.byref(params), _vp(film_ptr), _vp(rays_ptr), _vp(kinds_ptr),
                                             C.c_uint64(cap), C.c_uint32(stride), C.byref(count), _vp(stream)))
        return int(count.value)

    def resolve(self, film: np.ndarray) -> np.ndarray:
        film = np.ascontiguousarray(film, dtype=np.float32)
        rgb = np.empty(film.shape[:2] + (3,), dtype=np.float32)
        _check(lib.iile_film_resolve(self._h, _ptr(film), _ptr(rgb)))
        return rgb

    def write_image(self, path, rgb: np.ndarray):
        i = self.info()
        cb = i.croppedBounds
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        _check(lib.iile_write_image(os.fsencode(str(path)), _ptr(rgb), rgb.shape[1], rgb.shape[0],
                                    i.setup.xRes, i.setup.yRes, cb[0], cb[1]))

    # -- IISPTdIntegrator::RenderView ------------------------------------------
    def render_hemispheres(self, pos: np.ndarray, direction: np.ndarray, size=32, spp=1, out=None):
        """IISPTdIntegrator::RenderView for a batch; `out` = caller-owned (intensity [n,s,s,3], normals
        [n,s,s,3], distance [n,s,s]) float32 buffers to reuse (page-locked ones are read back by DMA)."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
        n = pos.shape[0]
        if out is not None:
            inten, nrm, dist = out
            for a, shape in ((inten, (n, size, size, 3)), (nrm, (n, size, size, 3)), (dist, (n, size, size))):
                if a.shape != shape or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError("out buffers must be C-contiguous float32 of shapes [n,s,s,3], [n,s,s,3], [n,s,s]")
        else:
            inten = np.zeros((n, size, size, 3), np.float32)
            nrm = np.zeros((n, size, size, 3), np.float32)
            dist = np.zeros((n, size, size), np.float32)
        stats = RenderStats()
        _check(lib.iile_render_hemispheres(self._h, n, _ptr(pos), _ptr(direction), size, spp,
                                           _ptr(inten), _ptr(nrm), _ptr(dist), C.byref(stats)))
        return inten, nrm, dist, stats


    def hemispheres_to_nn_dev(self, pos, direction, nn_ptr: int, means_ptr: int, size=32, spp=1, stream: int = 0,
                              want_stats: bool = False):
        """RenderView + IisptRenderRunner::normalizeMapsDownstream for a batch of hemispheres, left in
        HBM as the network's input: nn_ptr -> float32 [n,7,size,size], means_ptr -> float32 [n,3]."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
        stats = RenderStats() if want_stats else None
        _check(lib.iile_hemispheres_to_nn_dev(self._h, pos.shape[0], _ptr(pos), _ptr(direction), size, spp,
                                              _vp(nn_ptr), _vp(means_ptr),
                                              C.byref(stats) if want_stats else None, _vp(stream)))
        return stats


    # -- one task of IisptRenderRunner::run ----------------------------------------------------------
    def iispt_hemi_points(self, task: IisptTask):
        """Stage 1 of a task (run :252-325): (pos [n,3], dir [n,3], valid [n]) of the task's hemi points."""
        gw, gh = C.c_int32(), C.c_int32()
        _check(lib.iile_iispt_grid(C.byref(task), C.byref(gw), C.byref(gh)))
        n = gw.value * gh.value
        pos, direction = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        valid = np.zeros(n, np.uint8)
        _check(lib.iile_iispt_hemi_points(self._h, C.byref(task), _ptr(pos), _ptr(direction), _ptr(valid)))
        return pos, direction, valid

    def iispt_gather_dev(self, task: IisptTask, pos, direction, valid, nn_intensity_ptr: int, out_ptr: int,
                         stream: int = 0):
        """Stage 3 of a task (run :404-570): per film pixel {f_beta * L, sample weight} into out_ptr
        (device float32 [(y1-y0)*(x1-x0), 4]); nn_intensity_ptr: device float32 [valid.sum(), size, size, 3]."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
        valid = np.ascontiguousarray(valid, dtype=np.uint8)
        _check(lib.iile_iispt_gather_dev(self._h, C.byref(task), _ptr(pos), _ptr(direction), _ptr(valid),
                                         _vp(nn_intensity_ptr), _vp(out_ptr), _vp(stream)))

    def iispt_task(self, task: IisptTask, net, spp=1, device="cuda"):
        """One whole task with the network called IN PROCESS (replacing IisptNnConnector's stdio round trip
        per hemisphere, iisptnnconnector.cpp:36-148): hemi points -> RenderView + normalizeMapsDownstream
        (device) -> net([n,7,s,s] torch tensor in HBM) -> transformMapsUpstream (device) -> gather.
        `net`: any callable from a float32 CUDA tensor [n,7,s,s] to [n,3,s,s] (ml/iispt_net.py's IISPTNet).
        Returns (samples [H,W,4] numpy: f_beta * L and the sample weight, pos, dir, valid, films tensor)."""
        import torch
        size = task.hemiSize
        pos, direction, valid = self.iispt_hemi_points(task)
        keep = valid.astype(bool)
        n = int(keep.sum())
        h, w = task.y1 - task.y0, task.x1 - task.x0
        films = torch.zeros((max(n, 1), size, size, 3), dtype=torch.float32, device=device)
        out = torch.zeros((h * w, 4), dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream().cuda_stream
        if n:
            nn_in = torch.empty((n, 7, size, size), dtype=torch.float32, device=device)
            means = torch.empty((n, 3), dtype=torch.float32, device=device)
            self.hemispheres_to_nn_dev(pos[keep], direction[keep], nn_in.data_ptr(), means.data_ptr(), size, spp, stream)
            with torch.no_grad():
                nn_out = net(nn_in).contiguous().float()
            if tuple(nn_out.shape) != (n, 3, size, size):
                raise ValueError(f"net returned {tuple(nn_out.shape)}, expected {(n, 3, size, size)}")
            hemispheres_from_nn_dev(n, size, nn_out.data_ptr(), means.data_ptr(), films.data_ptr(), stream)
        self.iispt_gather_dev(task, pos, direction, valid, films.data_ptr(), out.data_ptr(), stream)
        torch.cuda.current_stream().synchronize()
        return out.cpu().numpy().reshape(h, w, 4), pos, direction, valid, films[:n]


    def iile_render(self, net, indirect_tasks=16, direct_samples=16, hemi_size=32, radius_start=100.0, seed=0,
                    hemi_spp=1):
        """IISPTIntegrator::render_normal_2 (iispt.cpp:357-447) with the device doing the work of the runner
        threads: IisptScheduleMonitor::next_task hands out `indirect_tasks` tasks (PbrtOptions.iileIndirectTasks), each
        runs through iispt_task (hemi points -> RenderView -> net in process -> gather) and its samples are added as
        film_monitor_indirect->add_n_samples adds them; the direct pass is `direct_samples` passes of
        DirectProgressiveIntegrator::RenderOnePass into film_monitor_direct (= Integrator "directlighting", maxdepth 5,
        one sample per pixel and pass, no pixel filter); the result is film_monitor_direct->merge_into(indirect): the sum
        of the two normalised films.  Returns (merged, direct, indirect) float32 [h, w, 3] over the film's bounds.

        The reference's film monitors cover Film::GetSampleBounds(); this mirror takes films whose sample bounds are the
        pixel bounds (pbrt's default box filter of half-width 0.5).  One runner: the sampler's pixel counter runs on
        from task to task; task k seeds its per-pixel gather streams with seed + (k << 32) (oracle: ref_harness
        iile_render, tests/test_iile_render_gpu.py)."""
        info = self.info()
        x0, y0, x1, y1 = [int(v) for v in info.sampleBounds]
        if [x0, y0, x1, y1] != [int(v) for v in info.croppedBounds]:
            raise IileError(E_UNSUPPORTED,
                            "iile_render: the film's sample bounds differ from its pixel bounds (a pixel filter wider "
                            "than the default box of half-width 0.5); the film monitors of this mirror cover the pixel bounds")
        w, h = x1 - x0, y1 - y0
        acc = np.zeros((h, w, 4), np.float64)  # IisptPixel: double r, g, b, weight
        schedule = IisptScheduleMonitor((x0, y0, x1, y1), radius_start)
        counter = 0
        while True:
            t = schedule.next_task()
            if t["taskNumber"] >= indirect_tasks:  # run :234-236
                break
            task = IisptTask(t["x0"], t["y0"], t["x1"], t["y1"], t["tilesize"], hemi_size, counter, 0,
                             (seed + (t["taskNumber"] << 32)) & 0xFFFFFFFFFFFFFFFF)
            samples = self.iispt_task(task, net, spp=hemi_spp)[0]
            gw, gh = C.c_int32(), C.c_int32()
            _check(lib.iile_iispt_grid(C.byref(task), C.byref(gw), C.byref(gh)))
            counter += gw.value * gh.value + (t["x1"] - t["x0"]) * (t["y1"] - t["y0"])
            took = samples[..., 3] > 0
            region = acc[t["y0"] - y0:t["y1"] - y0, t["x0"] - x0:t["x1"] - x0]
            region[..., :3] += np.where(took[..., None], samples[..., :3], 0).astype(np.float64)
            region[..., 3] += np.where(took, samples[..., 3], 0).astype(np.float64)
        indirect = np.where(acc[..., 3:4] > 0, acc[..., :3] / np.maximum(acc[..., 3:4], 1e-300), acc[..., :3]).astype(np.float32)
        # run_direct (iisptrenderrunner.cpp:601-631)
        keep = RenderSetup.from_buffer_copy(info.setup)
        s = RenderSetup.from_buffer_copy(info.setup)
        s.integratorType, s.directStrategy, s.maxDepth = 1, 0, 5  # DirectProgressiveIntegrator: UniformSampleAllLights, maxDepth 5
        s.pixelSamples, s.samplerType, s.hasPixelBounds = direct_samples, SAMPLER_RANDOM, 0
        s.filterType, s.filterRadius[0], s.filterRadius[1] = FILTER_MONITOR, 0.5, 0.5
        s.filmScale, s.maxSampleLuminance = 1.0, float("inf")
        self.set_setup(s)
        try:
            direct = self.resolve(self.render()[0])
        finally:
            self.set_setup(keep)
        return (direct + indirect).astype(np.float32), direct, indirect


class IisptScheduleMonitor:
    """IisptScheduleMonitor (iisptschedulemonitor.cpp:9-76): tasks of NUMBER_TILES x NUMBER_TILES hemi-point cells tile
    the film row by row; after every sweep the cell size ("radius") shrinks by sqrt(0.79541357).  radius_start /
    radius_ratio are the reference's IISPT_SCHEDULE_RADIUS_START / _RATIO environment variables."""
    NUMBER_TILES = 10

    def __init__(self, bounds, radius_start=100.0, radius_ratio=None):
        self.x0, self.y0, self.x1, self.y1 = [int(v) for v in bounds]
        self.radius = np.float32(radius_start)
        self.ratio = np.sqrt(np.float32(0.79541357)) if radius_ratio is None else np.float32(radius_ratio)
        self.nextx, self.nexty = self.x0, self.y0
        self.pass_, self.task_number = 1, 0

    def next_task(self):
        radius = max(int(np.floor(self.radius)), 1)
        size = radius * self.NUMBER_TILES
        t = {"x0": self.nextx, "y0": self.nexty, "x1": min(self.nextx + size, self.x1), "y1": min(self.nexty + size, self.y1),
             "tilesize": radius, "pass": self.pass_, "taskNumber": self.task_number}
        self.task_number += 1
        self.nextx += size
        if self.nextx >= self.x1:
            self.nextx = self.x0
            self.nexty += size
        if self.nexty >= self.y1:
            self.nexty = self.y0
            self.radius = np.float32(self.radius * self.ratio)
            self.pass_ += 1
        return t


def hemispheres_from_nn_dev(n: int, size: int, nn_out_ptr: int, means_ptr: int, intensity_ptr: int, stream: int = 0):
    """IisptRenderRunner::transformMapsUpstream on the device: network output [n,3,size,size] + the
    means of the matching to_nn call -> intensity [n,size,size,3] (ImageFilm row order)."""
    _check(lib.iile_hemispheres_from_nn_dev(n, size, _vp(nn_out_ptr), _vp(means_ptr), _vp(intensity_ptr), _vp(stream)))


def generate_rays_dev(bmin, bmax, first_ray: int, n: int, rays_ptr: int, stream: int = 0):
    a = (C.c_float * 3)(*[float(x) for x in bmin])
    b = (C.c_float * 3)(*[float(x) for x in bmax])
    _check(lib.iile_generate_rays_dev(a, b, first_ray, n, _vp(rays_ptr), _vp(stream)))


COMM_ID_BYTES = 128


class FilmComm:
    """NCCL communicator for the film merge (Film::MergeFilmTile across GPUs, film.cpp:135-148),
    made through the C ABI: rank 0 draws the unique id, `exchange(id_bytes) -> id_bytes` ships it
    to the other ranks (e.g. a torch.distributed broadcast_object_list, a file, a socket), every
    rank then joins on its current CUDA device."""

    def __init__(self, rank: int, world: int, exchange):
        self.rank, self.world, self._h = rank, world, _vp()
        ident = b""
        if rank == 0:
            buf = (C.c_uint8 * COMM_ID_BYTES)()
            _check(lib.iile_comm_unique_id(buf))
            ident = bytes(buf)
        ident = exchange(ident)
        if len(ident) != COMM_ID_BYTES:
            raise ValueError("FilmComm: the exchanged id must be %d bytes" % COMM_ID_BYTES)
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(ident)
        _check(lib.iile_comm_create(buf, world, rank, C.byref(self._h)))

    @staticmethod
    def nccl_version() -> int:
        v = C.c_int()
        _check(lib.iile_comm_nccl_version(C.byref(v)))
        return v.value

    @property
    def handle(self) -> int:
        return self._h.value or 0

    def reduce(self, film_ptr: int, n_pixels: int, root: int = 0, stream: int = 0):
        """In-place sum of the per-rank {R,G,B,w} films on `stream`; root < 0 = onto every rank."""
        _check(lib.iile_film_reduce_dev(_vp(film_ptr), n_pixels, root, self._h, _vp(stream)))

    def close(self):
        if self._h:
            lib.iile_comm_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


from . import partition  # noqa: E402  (multi-GPU work split + film merge)


def read_pfm(path) -> np.ndarray:
    """PFM reader for tests/tools (top-to-bottom rows, like Film's pixel order)."""
    with open(path, "rb") as f:
        kind = f.readline().strip()
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        ch = 3 if kind == b"PF" else 1
        a = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, ch)
    return a[::-1].astype(np.float32)
