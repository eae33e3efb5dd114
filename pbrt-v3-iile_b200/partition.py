"""Multi-GPU work partition of SamplerIntegrator::Render and the film merge.

The path shards by (pixel, sample): the scene is replicated on every GPU, rank r renders a
sample range (partition A) or a range of 16x16 tiles (partition B) of the same image through
``IileRenderParams``, and the only exchange is the sum of the per-rank ``{R,G,B,weight}`` films —
``Film::MergeFilmTile`` (src/core/film.cpp:135-148) across processes instead of across threads.
On GPUs the sum is ``iile_film_reduce_dev`` — one NCCL reduce over NVLink issued through the C ABI
on the render's own stream, with a communicator made by ``FilmComm`` (``film_comm`` below ships
the NCCL unique id over the process group the launcher set up).  Without a ``FilmComm`` the sum
is one ``torch.distributed`` reduce (what the gloo CPU tests exercise).  There is no other
collective on the path.

Nothing here computes pixels: a film tensor is whatever ``Scene.render_dev`` produced.
"""
from __future__ import annotations


def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    """Samples [begin, end) of every pixel for `rank`: contiguous, disjoint, covering [0, spp),
    sizes differing by at most one (the first spp % world ranks take one more)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    if spp < 0:
        raise ValueError("negative sample count")
    base, extra = divmod(spp, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def tile_range(n_tiles: int, rank: int, world: int) -> tuple[int, int]:
    """16x16 tiles [begin, end) in Render's tile order (integrator.cpp:237-245) for `rank`."""
    return sample_range(n_tiles, rank, world)


def tile_count(sample_bounds) -> int:
    """Number of 16x16 tiles covering Film::GetSampleBounds (x0, y0, x1, y1)."""
    x0, y0, x1, y1 = sample_bounds
    return ((x1 - x0 + 15) // 16) * ((y1 - y0 + 15) // 16)


def exchange_id(ident: bytes, src: int = 0, group=None) -> bytes:
    """Ship rank `src`'s NCCL unique id to every rank of the torch.distributed group (any backend)."""
    import torch.distributed as dist
    box = [ident]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def film_comm(rank: int, world: int, group=None):
    """A ``FilmComm`` over the ranks of the torch.distributed group (the launcher's rendezvous is
    only the side channel for the 128-byte id; the film bytes go through the C ABI's NCCL call)."""
    import sys
    FilmComm = sys.modules[__name__.rsplit(".", 1)[0]].FilmComm
    return FilmComm(rank, world, lambda ident: exchange_id(ident, 0, group))


def load_scene_shared(scene_cls, path, rank: int, world: int, *, upload: bool = True, flat_dir=None, group=None):
    """Every rank's copy of one scene with ONE parse + BVH build: rank 0 reads the .pbrt file,
    builds the accelerator on all host cores and writes the flat scene next to the ranks' scratch
    space; after a barrier the other ranks of the node load that file (no parse, no build — the
    nodes and the primitive order travel in it) and upload it to their own GPU.  `scene_cls` is the
    package's ``Scene``.  With world == 1 this is ``Scene.from_pbrt``."""
    if world <= 1:
        return scene_cls.from_pbrt(path, upload=upload)
    import os
    import shutil
    import tempfile
    import torch.distributed as dist
    # rank 0 makes a private directory (mode 0700, unpredictable name) and broadcasts its path: a
    # fixed name in a world-writable directory could be planted or raced by another local user
    box = [None]
    scene = None
    if rank == 0:
        box[0] = tempfile.mkdtemp(prefix="iile_shared_", dir=flat_dir)
        scene = scene_cls.from_pbrt(path, upload=upload)
        scene.save_flat(os.path.join(box[0], "scene.flat"))
    dist.broadcast_object_list(box, src=0, group=group)
    if rank != 0:
        scene = scene_cls.from_flat(os.path.join(box[0], "scene.flat"), upload=upload)
    dist.barrier(group=group)
    if rank == 0:
        shutil.rmtree(box[0], ignore_errors=True)
    return scene


def merge_films(film, dst: int | None = 0, group=None, comm=None, stream: int = 0):
    """Sum the per-rank films in place: onto `dst` (reduce) or onto every rank when dst is None
    (all-reduce).  `film` is a torch tensor [H, W, 4]; returns it.  With `comm` (a FilmComm) the
    sum is the C ABI's NCCL reduce on `stream`."""
    if comm is not None:
        comm.reduce(film.data_ptr(), film.shape[0] * film.shape[1], -1 if dst is None else dst, stream)
        return film
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return film
    if dst is None:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


def render_partitioned(scene, film, rank: int, world: int, *, by: str = "samples", stream: int = 0,
                       dst: int | None = 0, params_type=None, comm=None):
    """Render this rank's share of the scene's image into `film` (a CUDA tensor [H, W, 4]) and
    merge.  `by` is "samples" (partition A) or "tiles" (partition B)."""
    info = scene.info()
    RenderParams = params_type
    if RenderParams is None:
        import sys
        RenderParams = sys.modules[type(scene).__module__].RenderParams
    if by == "samples":
        b, e = sample_range(info.setup.pixelSamples, rank, world)
        params = RenderParams(b, e, 0, 0, 1, 0, 0)
        empty = b == e
    elif by == "tiles":
        b, e = tile_range(tile_count(tuple(info.sampleBounds)), rank, world)
        params = RenderParams(0, 0, b, e, 1, 0, 0)
        empty = b == e
    else:
        raise ValueError(f"unknown partition {by!r}")
    if empty:
        film.zero_()
    else:
        if by == "tiles" and e == 0:
            raise ValueError("empty tile range")
        scene.render_dev(film.data_ptr(), params, stream, want_stats=False)
    return merge_films(film, dst, comm=comm, stream=stream)
