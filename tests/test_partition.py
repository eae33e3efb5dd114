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
