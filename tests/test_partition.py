"""Multi-GPU host logic on CPU: the work partition (sample / tile ranges) and the film merge
over torch.distributed with the gloo backend, world_size 2.  No pixels are
computed here — films are synthetic — so no CUDA is needed."""
import os
import socket

import numpy as np
import pytest


def test_sample_ranges_cover_disjointly(iile):
    P = iile.partition
    for spp in (0, 1, 7, 16, 64, 255):
        for world in (1, 2, 3, 4, 8):
            ranges = [P.sample_range(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.sample_range(16, 2, 2)
    # cfg1: 400x400, box filter -> sample bounds [0,400)^2 -> 25x25 tiles (integrator.cpp:237-240)
    assert P.tile_count((0, 0, 400, 400)) == 625
    # 32x32 film with the Gaussian r=2 filter: sample bounds [-2,34)^2 -> 3x3 tiles
    assert P.tile_count((-2, -2, 34, 34)) == 9
    assert P.tile_range(625, 7, 8) == P.sample_range(625, 7, 8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, root, results):
    import importlib.util
    import sys
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("iile_partition", os.path.join(root, "pbrt-v3-iile_b200", "partition.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    # a synthetic per-rank film: this rank's sample range of a 4-spp, 8x6 image where sample s of
    # pixel p contributes (p + s, 2p, s, 1)
    h, w, spp = 6, 8, 5
    b, e = P.sample_range(spp, rank, world)
    pix = torch.arange(h * w, dtype=torch.float32).reshape(h, w)
    film = torch.zeros(h, w, 4)
    for s in range(b, e):
        film += torch.stack([pix + s, 2 * pix, torch.full_like(pix, float(s)), torch.ones_like(pix)], dim=-1)
    reduced = P.merge_films(film.clone(), dst=0)
    allred = P.merge_films(film.clone(), dst=None)
    want = torch.zeros(h, w, 4)
    for s in range(spp):
        want += torch.stack([pix + s, 2 * pix, torch.full_like(pix, float(s)), torch.ones_like(pix)], dim=-1)
    ok = bool(torch.equal(allred, want)) and (rank != 0 or bool(torch.equal(reduced, want)))
    # the side channel of FilmComm: rank 0's 128-byte NCCL id reaches every rank unchanged
    ident = bytes(range(128)) if rank == 0 else b""
    ok = ok and P.exchange_id(ident) == bytes(range(128))
    # one parse + BVH build per node: rank 0 flattens, rank 1 loads its flat scene (no CUDA: upload=False)
    sys.path.insert(0, os.path.join(root, "tests"))
    from conftest import load_product
    iile = load_product()
    scene = P.load_scene_shared(iile.Scene, os.path.join(root, "scenes", "smoke.pbrt"), rank, world, upload=False)
    whole = iile.Scene.from_pbrt(os.path.join(root, "scenes", "smoke.pbrt"), upload=False)
    ok = ok and scene.nodes().tobytes() == whole.nodes().tobytes()
    ok = ok and bool((scene.prim_order() == whole.prim_order()).all())
    ok = ok and (rank == 0) == (scene.info().buildSeconds > 0)      # only rank 0 built
    results[rank] = ok
    dist.destroy_process_group()


def test_film_merge_over_gloo_world2():
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, root, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(results) == {0: True, 1: True}


def test_comm_api_without_gpu(iile):
    """The NCCL film-reduce entry points: NCCL binds at run time, a NULL / absent
    communicator is the single-GPU no-op, and making a communicator needs a device (no CPU path)."""
    import ctypes as C
    assert iile.FilmComm.nccl_version() >= 20000
    assert iile.lib.iile_film_reduce_dev(None, 0, 0, None, None) == 0
    assert iile.lib.iile_film_reduce_dev(C.c_void_p(16), 4, 0, None, None) == 0   # comm NULL: untouched
    assert iile.lib.iile_film_reduce_dev(None, 4, 0, None, None) == -1
    assert iile.lib.iile_comm_destroy(None) == 0
    if iile.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(iile.IileError) as e:
        iile.FilmComm(0, 1, lambda ident: ident)
    assert e.value.code == -3
    with pytest.raises(ValueError):
        iile.FilmComm(1, 2, lambda ident: b"short")


def _nccl_worker(rank, world, port, root, results):
    import sys
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.join(root, "tests"))
    from conftest import load_product
    iile = load_product()
    torch.cuda.set_device(rank)
    iile.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # side channel only
    comm = iile.partition.film_comm(rank, world)
    scene = iile.Scene.from_flat(os.path.join(root, "tests", "golden", "smoke.flat"))
    setup = scene.info().setup
    setup.xRes, setup.yRes, setup.pixelSamples = 96, 64, 8
    scene.set_setup(setup)
    h, w = scene.film_shape()
    stream = torch.cuda.Stream()
    ok = True
    for by in ("samples", "tiles"):
        film = torch.zeros(h, w, 4, device="cuda")
        with torch.cuda.stream(stream):
            iile.partition.render_partitioned(scene, film, rank, world, by=by, stream=stream.cuda_stream,
                                              dst=None, comm=comm)
        stream.synchronize()
        whole = torch.zeros(h, w, 4, device="cuda")
        scene.render_dev(whole.data_ptr(), iile.RenderParams(0, 0, 0, 0, 1, 0, 0), 0)
        torch.cuda.synchronize()
        # same samples, same streams: only the float summation order of the film differs
        ok = ok and bool(torch.allclose(film, whole, rtol=2e-5, atol=1e-6))
        ok = ok and bool(torch.equal(film[..., 3], whole[..., 3]))     # box-filter weights are integers
    comm.close()
    results[rank] = ok
    dist.destroy_process_group()


@pytest.mark.gpu
def test_film_reduce_nccl_world2(iile):
    """Two GPUs, one process each: sample- and tile-partitioned renders merged by the C ABI's NCCL
    all-reduce equal the one-GPU film (needs >= 2 devices; the driver's -m gpu box may have one)."""
    if iile.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, root, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(results) == {0: True, 1: True}
