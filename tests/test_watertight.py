"""The reference's own known-answer test for the cast path, `Triangle.Watertight`
(src/tests/shapes.cpp:28-152), restated against the C ABI: a triangulated sphere whose vertices are
pushed out along their normals by random amounts (poles and the seam shared exactly), rays from
random points inside it in random directions and — "now tougher" — straight at mesh vertices.  Every
ray must hit: no ray may slip between two triangles that share an edge or a vertex.

The mesh is handed over through iile_scene_create (caller arrays), the CPU half checks the oracle
restatement, the GPU half the kernels (closest-hit and any-hit) and their bit-equality with it."""
import numpy as np
import pytest

from conftest import ROOT

GOLD = ROOT / "tests" / "golden"
N_THETA = N_PHI = 16


def sphere_mesh(seed=12111):
    rng = np.random.default_rng(seed)
    verts = []
    for t in range(N_THETA):
        theta = np.float32(np.pi) * np.float32(t) / np.float32(N_THETA - 1)
        ct, st = np.cos(theta, dtype=np.float32), np.sin(theta, dtype=np.float32)
        for p in range(N_PHI):
            phi = np.float32(2 * np.pi) * np.float32(p) / np.float32(N_PHI - 1)
            if t == 0:
                verts.append((0, 0, 1))
            elif t == N_THETA - 1:
                verts.append((0, 0, -1))
            elif p == N_PHI - 1:
                verts.append(verts[len(verts) - (N_PHI - 1)])      # close it up exactly at the end
            else:
                r = np.float32(1 + 5 * rng.random())
                verts.append((r * st * np.cos(phi), r * st * np.sin(phi), r * ct))
    P = np.asarray(verts, np.float32)
    off = lambda t, p: t * N_PHI + p
    idx = []
    for p in range(N_PHI - 1):                                      # fan at top
        idx.append((off(0, 0), off(1, p), off(1, p + 1)))
    for t in range(1, N_THETA - 2):                                 # quads in the middle rows
        for p in range(N_PHI - 1):
            idx.append((off(t, p), off(t + 1, p), off(t + 1, p + 1)))
            idx.append((off(t, p), off(t + 1, p + 1), off(t, p + 1)))
    for p in range(N_PHI - 1):                                      # fan at bottom
        idx.append((off(N_THETA - 1, 0), off(N_THETA - 2, p), off(N_THETA - 2, p + 1)))
    return P, np.asarray(idx, np.int32)


def uniform_sphere(u0, u1):
    z = 1 - 2 * u0
    r = np.sqrt(np.maximum(0, 1 - z * z))
    phi = 2 * np.pi * u1
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=-1).astype(np.float32)


def rays_for(iile, P, n=100000, seed=7):
    rng = np.random.default_rng(seed)
    o = np.float32(0.5) * uniform_sphere(rng.random(n), rng.random(n))
    rays = np.zeros(2 * n, dtype=iile.RAY_DTYPE)
    rays["o"][:n] = rays["o"][n:] = o
    rays["d"][:n] = uniform_sphere(rng.random(n), rng.random(n))
    rays["d"][n:] = P[rng.integers(0, len(P), n)] - o               # straight at a vertex (unnormalised, as the reference)
    rays["tMax"] = np.inf
    return rays


def make_scene(iile, upload, copies=1):
    P, idx = sphere_mesh()
    idx = np.concatenate([idx] * copies)                            # copies > 1: coincident triangles, exact ties
    prims = np.zeros(len(idx), dtype=iile.PRIM_DTYPE)
    prims["type"], prims["v"], prims["material"], prims["light"] = 0, idx, 0, -1
    setup = iile.Scene.from_flat(GOLD / "smoke.flat", upload=False).info().setup
    mats = np.zeros((1, iile.MATERIAL_BYTES), np.uint8)             # one matte material, Kd = 0
    return iile.Scene.from_desc(P, prims, setup, materials=mats, upload=upload), P, len(idx)


def test_watertight_oracle(iile, oracle, tmp_path):
    scene, P, n_tris = make_scene(iile, upload=False)
    assert n_tris == 2 * (N_PHI - 1) + 2 * (N_THETA - 3) * (N_PHI - 1)
    scene.save_flat(tmp_path / "w.flat")
    rays = rays_for(iile, P, n=20000)
    hits, _, _ = oracle.cast(oracle.FlatScene(tmp_path / "w.flat"), rays)
    assert (hits["prim"] >= 0).all(), "the restatement lets a ray through the mesh"


@pytest.mark.gpu
def test_watertight_gpu(iile, oracle, tmp_path):
    if iile.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    iile.set_device(0)
    scene, P, _ = make_scene(iile, upload=True)
    rays = rays_for(iile, P)                                         # 100,000 + 100,000 rays, as the reference
    hits = scene.intersect(rays)
    assert (hits["prim"] >= 0).all(), f"{int((hits['prim'] < 0).sum())} rays leaked through the mesh"
    assert scene.intersect_p(rays).all()
    scene.save_flat(tmp_path / "w.flat")
    want, _, _ = oracle.cast(oracle.FlatScene(tmp_path / "w.flat"), rays)
    assert np.array_equal(hits["prim"], want["prim"]) and np.array_equal(hits["t"].view("u4"), want["t"].view("u4"))


@pytest.mark.gpu
def test_coincident_triangles_resolve_like_the_reference(iile, oracle, tmp_path):
    """Every triangle twice: each hit is an exact tie between two primitives.  The reference keeps the
    one its traversal reaches last (Triangle::Intersect accepts t == ray.tMax, triangle.cpp:247-261);
    the kernels must name the same primitive, bit for bit, as the restatement of that traversal."""
    if iile.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    iile.set_device(0)
    scene, P, n_tris = make_scene(iile, upload=True, copies=2)
    rays = rays_for(iile, P, n=50000)
    hits = scene.intersect(rays)
    scene.save_flat(tmp_path / "w2.flat")
    want, _, _ = oracle.cast(oracle.FlatScene(tmp_path / "w2.flat"), rays)
    assert (want["prim"] >= 0).all()
    assert np.array_equal(hits["prim"], want["prim"]) and np.array_equal(hits["t"].view("u4"), want["t"].view("u4"))
    both = np.bincount(hits["prim"] // (n_tris // 2), minlength=2)
    assert both.min() > 0                                           # both copies win somewhere: the order matters
