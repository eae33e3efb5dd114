"""BASELINE configs[3] — the IISPT hemispheric auxiliary pass (IISPTdIntegrator::RenderView,
src/integrators/iispt_d.cpp:345-461) batched on the GPU through iile_render_hemispheres.

Oracle: the unmodified reference's HemisphericCamera + IISPTdIntegrator (ref_harness hemi) with
the same per-(pixel,sample) streams; committed as tests/golden/hemi_ref.npz and re-run on the
box when oracle/_ref travelled.  Bars: miss flags identical; first-hit distances and camera-space normals BIT-identical
(the camera direction goes through sinf / cosf, which carry the host libm's bits on the device); intensity — a sum of path
contributions in another order — >= 97 % of the pixels within 1e-5 relative, >= 99.99 % within 1e-3, relMSE <= 1e-9."""
import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
GOLD = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def gpu(iile):
    if iile.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    iile.set_device(0)
    return iile


def check(inten, nrm, dist, want_i, want_n, want_z):
    assert np.array_equal(dist < 0, want_z < 0), "hit / miss pattern differs"
    hit = want_z >= 0
    assert np.all(dist[~hit] == -1.0)
    rel = np.abs(dist[hit] - want_z[hit]) / want_z[hit]
    # the camera direction goes through std::sin / std::cos of floats: with the host libm's bits on the device
    # (csrc/cuda/libm_exact.cuh) the hemisphere's rays are the reference's, so first-hit distances and normals are
    # BIT-identical (they agreed to 1e-5 / 2e-5 while CUDA's sinf / cosf were in the way)
    assert np.array_equal(dist.view("u4"), want_z.view("u4")), "first-hit distances differ"
    assert np.array_equal(nrm.view("u4"), want_n.view("u4")), "first-hit normals differ"
    d = np.abs(inten - want_i)
    frac = float(((d / np.maximum(np.abs(want_i), 1e-3)).max(axis=-1) < 1e-3).mean())
    relmse = float((((inten - want_i) ** 2).mean(axis=-1) / (want_i.mean(axis=-1) ** 2 + 1e-2)).mean())
    f5 = float(((d / np.maximum(np.abs(want_i), 1e-3)).max(axis=-1) < 1e-5).mean())
    print(f"HEMISTAT distance bit-equal {np.array_equal(dist.view('u4'), want_z.view('u4'))} (max rel {rel.max(initial=0):.3g})  "
          f"normals bit-equal {np.array_equal(nrm.view('u4'), want_n.view('u4'))} (max abs {np.abs(nrm - want_n).max():.3g})  "
          f"intensity within 1e-5: {f5:.4f} within 1e-3: {frac:.4f} relMSE {relmse:.3g}")
    # intensity: a sum of path contributions in another order — measured 98.9 ... 100 % of the pixels within 1e-5,
    # every pixel within 1e-3, relMSE <= 2e-14 (profiles/r10_hemi_iispt_tests.log)
    assert f5 >= 0.97 and frac >= 0.9999 and relmse <= 1e-9, (f5, frac, relmse)
    return frac, relmse


def test_hemispheres_match_golden(gpu):
    g = np.load(GOLD / "hemi_ref.npz")
    scene = gpu.Scene.from_flat(GOLD / "smoke.flat")
    inten, nrm, dist, stats = scene.render_hemispheres(g["pos"], g["dir"], size=32, spp=1)
    assert inten.shape == (8, 32, 32, 3) and stats.cameraRays == 8 * 32 * 32
    check(inten, nrm, dist, g["intensity"], g["normals"], g["distance"])
    # a batch is independent of how it is cut: hemisphere h keeps stream index h only when it
    # keeps its place, so compare a prefix
    i2, n2, z2, _ = scene.render_hemispheres(g["pos"][:3], g["dir"][:3], size=32, spp=1)
    assert np.array_equal(z2, dist[:3]) and np.array_equal(n2, nrm[:3])
    assert np.allclose(i2, inten[:3], rtol=2e-5, atol=1e-7)


def test_hemispheres_cfg1_vs_reference_on_this_box(gpu, oracle, cfg1_flat, workdir):
    """Hemispheres at the first-hit points of a 6x6 grid of cfg1 camera pixels."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    scene = gpu.Scene.from_flat(cfg1_flat)
    flat = oracle.FlatScene(cfg1_flat)
    lo, hi = oracle.killeroo_ray_box(flat)
    rays = oracle.generate_rays(lo, hi, 0, 4096)
    ref, _ = oracle.ref_raycast(cfg1_flat, rays, workdir)
    hit = np.flatnonzero(ref["prim"] >= 0)[:36]
    n = ref["n"][hit].astype(np.float32)
    n = np.where((n * rays["d"][hit]).sum(axis=1, keepdims=True) > 0, -n, n)
    pos = (ref["p"][hit] + np.float32(0.05) * n).astype(np.float32)
    want_i, want_n, want_z, _ = oracle.ref_hemispheres(cfg1_flat, pos, n, workdir)
    inten, nrm, dist, stats = scene.render_hemispheres(pos, n)
    check(inten, nrm, dist, want_i, want_n, want_z)
    assert stats.closestRays > 36 * 1024
    # several samples per hemisphere pixel: the distance and normal maps keep the LAST sample's first hit
    # (DistanceFilm / NormalFilm::set_camera_coord overwrite, iispt_d.cpp:98-112)
    want_i, want_n, want_z, _ = oracle.ref_hemispheres(cfg1_flat, pos[:8], n[:8], workdir, spp=3)
    inten, nrm, dist, _ = scene.render_hemispheres(pos[:8], n[:8], spp=3)
    check(inten, nrm, dist, want_i, want_n, want_z)


def test_hemisphere_argument_errors(gpu):
    scene = gpu.Scene.from_flat(GOLD / "smoke.flat")
    with pytest.raises(gpu.IileError):
        scene.render_hemispheres(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32), size=0)
    i, n, z, _ = scene.render_hemispheres(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert i.shape == (0, 32, 32, 3)


def _nn_check(nn, means, want_nn, want_means, raw):
    """nn [n,7,s,s] vs the reference's normalizeMapsDownstream output.  The maps inherit the bars of
    `check`; the transform itself is checked by
    re-deriving it from raw maps of the same batch (`raw` = intensity, normals, distance of a second
    render: normals and distance are bit-stable, the film sum is a float atomicAdd whose order — and
    so last bit — varies between runs, hence 5e-7 absolute on the intensity planes)."""
    inten, nrm, dist = raw
    # exact: planes 3..5 are the clamped camera-space normals, rows flipped (ImageFilm order)
    assert np.array_equal(nn[:, 3:6], np.clip(nrm, -1, 1)[:, ::-1].transpose(0, 3, 1, 2))
    n, s = len(nn), nn.shape[-1]
    for h in range(n):
        # serial double sums in ImageFilm order, as imagefilm.cpp:203-254
        flip = inten[h, ::-1].reshape(-1, 3)
        rs = gs = bs = tot = 0.0
        for r, g, b in flip:
            rs += float(r); gs += float(g); bs += float(b); tot += float(np.float32(np.float32(r + g) + b))
        m = np.array([rs / len(flip), gs / len(flip), bs / len(flip)], dtype=np.float32)
        assert np.allclose(means[h], m, rtol=1e-6, atol=1e-9)
        imean = np.float32(tot / (3 * len(flip)))
        ratio = np.float32(0.0) if imean == 0 else np.float32(1.0 / (10.0 * float(imean)))
        v = np.maximum(inten[h, ::-1] * ratio, np.float32(0))
        want = (np.log(1.0 + v.astype(np.float64)).astype(np.float32) + np.float32(-0.1)).transpose(2, 0, 1)
        assert np.abs(nn[h, :3] - want).max() <= 5e-7
        zs = 0.0
        for z in dist[h, ::-1].reshape(-1):
            zs += float(z)
        zmean = np.float32(zs / (s * s))
        div = np.float32(10.0 * (float(zmean) + 1.0))
        if div == 0:
            div = np.float32(1.0)
        zr = np.float32(1.0 / float(div))
        z = np.maximum((dist[h, ::-1] + np.float32(1.0)) * zr, np.float32(0))
        wantz = np.log(1.0 + z.astype(np.float64)).astype(np.float32) + np.float32(-0.1)
        assert np.abs(nn[h, 6] - wantz).max() <= 1.5e-8      # distance is bit-stable: CUDA log vs libm only
    if want_nn is None:
        return
    # against the reference: same statistics bars as the raw maps
    assert np.allclose(means, want_means, rtol=2e-3, atol=1e-6)
    assert np.array_equal(nn[:, 6] < np.float32(-0.05), want_nn[:, 6] < np.float32(-0.05))   # all-miss maps agree
    d = np.abs(nn - want_nn)
    assert np.mean(d[:, 3:6] == 0) > 0.98 and d[:, 3:6].max() <= 2e-5
    assert np.mean(d[:, 6] <= 1e-5) >= 0.995
    assert np.mean(d[:, :3] <= 2e-3) >= 0.95


def test_nn_maps_match_the_reference_runner(gpu):
    """iile_hemispheres_to_nn_dev / iile_hemispheres_from_nn_dev vs IisptRenderRunner's own
    normalizeMapsDownstream / transformMapsUpstream (golden: ref_harness hemi_nn on smoke.flat)."""
    import torch
    g = np.load(GOLD / "hemi_ref.npz")
    gn = np.load(GOLD / "hemi_nn_ref.npz")
    scene = gpu.Scene.from_flat(GOLD / "smoke.flat")
    n, s = len(g["pos"]), 32
    nn = torch.empty(n, 7, s, s, device="cuda")
    means = torch.empty(n, 3, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        stats = scene.hemispheres_to_nn_dev(g["pos"], g["dir"], nn.data_ptr(), means.data_ptr(), size=s, spp=1,
                                            stream=stream.cuda_stream, want_stats=True)
    stream.synchronize()
    assert stats.cameraRays == n * s * s
    raw = scene.render_hemispheres(g["pos"], g["dir"], size=s, spp=1)[:3]
    _nn_check(nn.cpu().numpy(), means.cpu().numpy(), gn["nn_input"], gn["means"], raw)
    # upstream: the golden's own normalised intensity as the network's answer (identity network) ->
    # pins the arithmetic of transformMapsUpstream.  std::exp(float) is glibc's expf, which is not
    # always correctly rounded (2 of these 24576 inputs differ from exp-in-double rounded once); one
    # ulp of e = exp(y) ~ 1 becomes ~1e-6 relative in e - 1, hence the two-level bar
    ident = torch.from_numpy(np.ascontiguousarray(gn["nn_input"][:, :3])).cuda()
    gm = torch.from_numpy(gn["means"]).cuda()
    out = torch.empty(n, s, s, 3, device="cuda")
    gpu.hemispheres_from_nn_dev(n, s, ident.data_ptr(), gm.data_ptr(), out.data_ptr(), 0)
    torch.cuda.synchronize()
    got, want = out.cpu().numpy(), gn["upstream"]
    # (measured on B200: 2 elements off by one ulp of e times the channel scale, everything else bit-identical)
    assert np.abs(got - want).max() <= 2e-6
    assert np.mean(got == want) > 0.999
    # all-black hemispheres (mean 0) stay exactly zero, as rmul = 0 (iisptrenderrunner.cpp:1114-1132)
    black = np.flatnonzero(gn["means"].max(axis=1) == 0)
    assert len(black) and np.all(got[black] == 0)
    # round trip of this run's own maps: upstream(downstream(x)) restores the channel means
    gpu.hemispheres_from_nn_dev(n, s, nn[:, :3].contiguous().data_ptr(), means.data_ptr(), out.data_ptr(), 0)
    torch.cuda.synchronize()
    back = out.cpu().numpy().reshape(n, -1, 3).mean(axis=1)
    assert np.allclose(back, means.cpu().numpy(), rtol=1e-5, atol=1e-8)


def test_nn_maps_argument_errors(gpu):
    scene = gpu.Scene.from_flat(GOLD / "smoke.flat")
    with pytest.raises(gpu.IileError):
        scene.hemispheres_to_nn_dev(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32), 0, 0)
    with pytest.raises(gpu.IileError):
        gpu.hemispheres_from_nn_dev(1, 32, 0, 0, 0)
    gpu.hemispheres_from_nn_dev(0, 32, 0, 0, 0)
    assert scene.hemispheres_to_nn_dev(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), 0, 0) is None


@pytest.mark.parametrize("size", [16, 64])
def test_nn_maps_other_hemisphere_sizes(gpu, size):
    """PbrtOptions.iisptHemiSize other than 32: 16 (planes staged in shared memory) and 64 (64 KB of
    planes: the kernel variant that sums from global memory), against the transform re-derived from the
    raw maps of the same batch."""
    import torch
    g = np.load(GOLD / "hemi_ref.npz")
    scene = gpu.Scene.from_flat(GOLD / "smoke.flat")
    pos, axis = g["pos"][:4], g["dir"][:4]
    nn = torch.empty(4, 7, size, size, device="cuda")
    means = torch.empty(4, 3, device="cuda")
    scene.hemispheres_to_nn_dev(pos, axis, nn.data_ptr(), means.data_ptr(), size=size, spp=1)
    torch.cuda.synchronize()
    raw = scene.render_hemispheres(pos, axis, size=size, spp=1)[:3]
    _nn_check(nn.cpu().numpy(), means.cpu().numpy(), None, None, raw)
