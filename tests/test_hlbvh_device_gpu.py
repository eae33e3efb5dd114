"""splitmethod "hlbvh" built on the GPU (csrc/cuda/hlbvh_device.cu; reference: BVHAccel::HLBVHBuild, emitLBVH,
buildUpperSAH, flattenBVHTree — src/accelerators/bvh.cpp:404-658).

Bars (integer / byte work: bit-exact):
  * the device build's LinearBVHNode array and primitive order are byte-identical to the host builder's
    (IILE_DEVICE_HLBVH=0), which tests/test_oracle_pinned.py pins against the reference built single-threaded;
  * where oracle/_ref travelled, also directly against BVHAccel's own arrays (axis / pad bytes of leaves are
    uninitialised there and are left out, as in the pinned test).
"""
import subprocess
import sys
import time

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(iile):
    if iile.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    iile.set_device(0)
    return iile


def build_both(gpu, scene_file, monkeypatch, capfd=None):
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("IILE_DEVICE_HLBVH", mode)
        monkeypatch.setenv("IILE_BUILD_TRACE", "1")
        t = time.time()
        s = gpu.Scene.from_pbrt(scene_file, upload=False)
        dt = time.time() - t
        assert s.info().setup.splitMethod == 1                   # IILE_SPLIT_HLBVH
        out[mode] = (s.nodes().copy(), s.prim_order().copy(), s.info().buildSeconds, dt, s)
    return out


def with_accel(ref_scenes, name, leaf):
    text = (ref_scenes / "cfg1.pbrt").read_text()
    assert "WorldBegin" in text and "Accelerator" not in text
    text = text.replace("WorldBegin", 'Accelerator "bvh" "string splitmethod" "hlbvh" "integer maxnodeprims" [%d]\nWorldBegin' % leaf, 1)
    f = ref_scenes / name                                        # next to cfg1.pbrt: it Includes geometry/killeroo.pbrt
    f.write_text(text)
    return f


@pytest.mark.parametrize("leaf", [4, 1, 9, 255])
def test_device_hlbvh_equals_host_build(gpu, oracle, ref_scenes, workdir, leaf, monkeypatch, capfd):
    """cfg1's killeroo (66,533 primitives incl. the sphere light, i.e. a quadric whose bounds come from the
    host) with maxnodeprims 1 / 4 / 9 / 255."""
    scene_file = with_accel(ref_scenes, f"_cfg1_dev_hlbvh_{leaf}.pbrt", leaf)
    try:
        got = build_both(gpu, scene_file, monkeypatch)
    finally:
        scene_file.unlink()
    err = capfd.readouterr().err
    assert "HLBVH on the device" in err, err                     # the device path really ran
    host, dev = got["0"], got["1"]
    assert len(host[0]) == len(dev[0])
    assert dev[0].tobytes() == host[0].tobytes(), "LinearBVHNode bytes differ between the device and the host build"
    assert np.array_equal(dev[1], host[1])
    if oracle.have_ref():
        flat = workdir / f"cfg1_dev_hlbvh_{leaf}.flat"
        dev[4].save_flat(flat)
        nodes, order, _ = oracle.ref_bvh(flat, workdir)
        mine = dev[0]
        for n in [n for n in nodes.dtype.names if n not in ("axis", "pad")]:
            assert np.array_equal(mine[n], nodes[n]), n
        interior = nodes["nPrims"] == 0
        assert np.array_equal(mine["axis"][interior], nodes["axis"][interior])
        assert np.array_equal(dev[1], order)
    # and the tree is usable: casts against it agree with the restatement walking the same arrays
    flat = workdir / f"cfg1_dev_hlbvh_{leaf}_cast.flat"
    dev[4].save_flat(flat)
    fs = oracle.FlatScene(flat)
    lo, hi = oracle.cfg2_ray_box(fs)
    rays = oracle.generate_rays(lo, hi, 3, 1 << 16)
    want, _, _ = oracle.cast(fs, rays)
    hits = dev[4].intersect(rays)
    assert np.array_equal(hits["prim"], want["prim"])
    assert np.array_equal(hits["t"].view("u4"), want["t"].view("u4"))


def test_device_hlbvh_coincident_cluster(gpu, tmp_path, monkeypatch, capfd):
    """Many primitives in one Morton cell (bvh.cpp:519: bit == -1 => one leaf of all of them) plus 70,000
    scattered triangles so that the scene is large enough for the device path."""
    rng = np.random.default_rng(5)
    verts, idx = [], []
    for i in range(400):                                         # 400 triangles around one centroid
        a = rng.random() * 2 * np.pi
        ca, sa = np.cos(a) * 1e-4, np.sin(a) * 1e-4
        verts += [(5 - ca, 5 - sa, 5.0), (5 + ca, 5 + sa, 5.0), (5.0, 5.0, 5.0)]
        idx += [3 * i, 3 * i + 1, 3 * i + 2]
    base = len(verts)
    pts = rng.random((70000, 3)) * 10
    for i, p in enumerate(pts):
        d = rng.random((2, 3)) * 0.05
        verts += [tuple(p), tuple(p + d[0]), tuple(p + d[1])]
        idx += [base + 3 * i, base + 3 * i + 1, base + 3 * i + 2]
    f = tmp_path / "cluster.pbrt"
    f.write_text('Accelerator "bvh" "string splitmethod" ["hlbvh"]\nSampler "random"\nWorldBegin\nMaterial "matte"\n'
                 'Shape "trianglemesh" "point P" [%s] "integer indices" [%s]\nWorldEnd\n' % (
                     " ".join("%.9g %.9g %.9g" % v for v in verts), " ".join(map(str, idx))))
    got = build_both(gpu, f, monkeypatch)
    assert "HLBVH on the device" in capfd.readouterr().err
    host, dev = got["0"], got["1"]
    assert host[0]["nPrims"].max() >= 256
    assert dev[0].tobytes() == host[0].tobytes() and np.array_equal(dev[1], host[1])


@pytest.mark.parametrize("lonely", [1, 2])
def test_device_hlbvh_lone_primitives_in_the_first_treelets(gpu, tmp_path, monkeypatch, capfd, lonely):
    """One (or two) far-away spheres whose Morton prefixes sort before everything else: treelets of a single primitive
    at sorted positions 0 and 1.  The first version of the device build marked such a treelet as ~position, which is
    the "no primitive" marker for position 0 (the treelet vanished: two nodes fewer than the reference) and the
    "internal error" marker for position 1 (found by tools/fuzz_bvh.py)."""
    rng = np.random.default_rng(11)
    pts = rng.random((70000, 3)) * 2 + 8                          # the soup sits in [8, 10]^3
    verts, idx = [], []
    for i, p in enumerate(pts):
        d = rng.random((2, 3)) * 0.02
        verts += [tuple(p), tuple(p + d[0]), tuple(p + d[1])]
        idx += [3 * i, 3 * i + 1, 3 * i + 2]
    spheres = 'AttributeBegin\nTranslate 0 0 0\nShape "sphere" "float radius" [0.1]\nAttributeEnd\n'
    if lonely == 2:  # a second one further along x: another prefix, still before the soup (its y and z bits are 0)
        spheres += 'AttributeBegin\nTranslate 3 0 0\nShape "sphere" "float radius" [0.1]\nAttributeEnd\n'
    f = tmp_path / "lonely.pbrt"
    f.write_text('Accelerator "bvh" "string splitmethod" ["hlbvh"]\nSampler "random"\nWorldBegin\nMaterial "matte"\n%s'
                 'Shape "trianglemesh" "point P" [%s] "integer indices" [%s]\nWorldEnd\n' % (
                     spheres, " ".join("%.9g %.9g %.9g" % v for v in verts), " ".join(map(str, idx))))
    got = build_both(gpu, f, monkeypatch)
    err = capfd.readouterr().err
    assert "HLBVH on the device" in err and "not used" not in err, err
    host, dev = got["0"], got["1"]
    order = host[1]
    assert list(order[:lonely]) == list(range(lonely))            # the spheres are primitives 0 (and 1) and sort first
    assert len(dev[0]) == len(host[0])
    assert dev[0].tobytes() == host[0].tobytes() and np.array_equal(dev[1], host[1])


@pytest.fixture(scope="module")
def cfg5_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg5h")
    subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_scenes.py"), "cfg5", str(d)],
                          stdout=subprocess.DEVNULL)
    return d


@pytest.mark.slow
def test_device_hlbvh_cfg5_terrain(gpu, cfg5_dir, monkeypatch, capfd):
    """BASELINE configs[4]'s 8,000,002 primitives: same bytes as the host HLBVH build, and the build itself
    several times faster than the host cores manage."""
    text = (cfg5_dir / "cfg5.pbrt").read_text()
    f = cfg5_dir / "cfg5_hlbvh.pbrt"
    f.write_text(text.replace("WorldBegin", 'Accelerator "bvh" "string splitmethod" "hlbvh"\nWorldBegin', 1))
    got = build_both(gpu, f, monkeypatch)
    err = capfd.readouterr().err
    assert "HLBVH on the device" in err
    host, dev = got["0"], got["1"]
    print(f"cfg5 hlbvh: host build {host[2]:.3f} s (load {host[3]:.2f} s), device build {dev[2]:.3f} s (load {dev[3]:.2f} s), "
          f"{len(dev[0])} nodes")
    print(err)
    assert len(dev[0]) == len(host[0]) > 2_000_000
    assert dev[0].tobytes() == host[0].tobytes()
    assert np.array_equal(dev[1], host[1])
    # the build proper (upload .. download, IILE_BUILD_TRACE's phases) against the host cores; the wall time of
    # the call also carries one cudaMalloc / cudaFree of the 1.6 GB arena, which depends on the process's state
    import re
    phases = [float(x) for x in re.findall(r"hlbvh\(device\): [a-z+ A-Z]+ ([0-9.]+) s", err)]
    assert len(phases) == 6 and sum(phases) < host[2], (phases, host[2])
