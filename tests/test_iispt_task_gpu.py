"""One task of IisptRenderRunner::run on the device (csrc/cuda/iispt_task.cu):
hemi points -> RenderView + normalizeMapsDownstream -> network IN PROCESS -> transformMapsUpstream ->
compute_fpixel_weights + sample_hemisphere / estimate_direct (iisptrenderrunner.cpp:16-178, 216-596, 939-1019).

Oracle: `ref_harness iispt_task` — the UNMODIFIED reference's own find_intersection, compute_fpixel_weights,
sample_hemisphere (hence estimate_direct, get_light_sample_nn, getLightSampleNn), run on this box with the same
per-sample / per-pixel PCG32 streams (oracle/ref_harness/iispt_task.cpp says what the harness replaces).

Bars (floating point; the tolerance is written here):
  * hemi points: validity identical, position and axis bit-identical (same arithmetic as the path integrator's
    SpawnRay, which is pinned bit-exact)
  * gather with THE SAME films on both sides (isolates the gather): which pixels add a sample is identical;
    >= 99.5 % of the samples within 1e-4 relative (measured: 100 %), task mean within 1e-4
  * behind specular bounces (the mirror / glass variants) the same bars: the hemi points behind a mirror or glass
    chain are bit-identical too since the device's sinf / cosf / acosf / atan2f carry the host libm's bits
    (csrc/cuda/libm_exact.cuh) — they were ">= 90 % within float precision" before
  * whole chain against the reference's own RenderView + identity network: task mean within 3 %
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZE = 32


@pytest.fixture(scope="module")
def gpu(iile):
    if iile.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    iile.set_device(0)
    return iile


def identity_net(x):
    """The network that answers with the (normalised) intensity it was given: enough to pin the arithmetic
    either side of it (as tests/test_hemi_gpu.py does for the map transforms)."""
    return x[:, 0:3]


def scene_variant(ref_scenes, name):
    text = (ref_scenes / "cfg1.pbrt").read_text()
    if name == "cfg1":
        return ref_scenes / "cfg1.pbrt", None
    assert 'Material "plastic"' in text
    repl = {"mirror": 'Material "mirror" "color Kr" [.9 .9 .9]',
            "glass": 'Material "glass" "float index" [1.5]'}[name]
    import re
    text = re.sub(r'Material "plastic"[^\n]*(\n\s+"[^\n]*)*', repl, text)
    f = ref_scenes / f"_cfg1_iispt_{name}.pbrt"
    f.write_text(text)
    return f, f


@pytest.mark.parametrize("variant,rect", [("cfg1", (150, 150, 214, 214, 16)), ("cfg1", (96, 200, 160, 233, 10)),
                                          ("mirror", (150, 150, 214, 214, 16)), ("glass", (150, 150, 214, 214, 16))])
def test_iispt_task_matches_reference(gpu, oracle, ref_scenes, workdir, variant, rect):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    import torch
    path, tmp = scene_variant(ref_scenes, variant)
    try:
        scene = gpu.Scene.from_pbrt(path)
    finally:
        if tmp:
            tmp.unlink()
    flat = workdir / f"iispt_{variant}.flat"
    scene.save_flat(flat)
    x0, y0, x1, y1, ts = rect
    task = gpu.IisptTask(x0, y0, x1, y1, ts, SIZE, 1000, 3, 424242)
    samples, pos, direction, valid, films = scene.iispt_task(task, identity_net)
    films_np = films.cpu().numpy()
    torch.cuda.synchronize()

    # ---- A: the reference walks the task with the device's films ------------------------------------
    hemi, _, pixels, info = oracle.ref_iispt_task(flat, workdir, x0, y0, x1, y1, ts, SIZE, 1000, 3, 424242, films=films_np)
    assert info["hemi_points"] == len(valid) and info["valid"] == int(valid.sum())
    assert np.array_equal(hemi[:, 0] > 0, valid.astype(bool)), "hemi points: validity differs"
    assert valid.sum() >= len(valid) // 2
    keep = valid.astype(bool)
    specular = variant != "cfg1"
    print(f"IISPTSTAT {variant} {rect}: hemi positions bit-equal {(hemi[keep, 1:4].view('u4') == pos[keep].view('u4')).all(axis=1).mean():.4f}  "
          f"axes bit-equal {(hemi[keep, 4:7].view('u4') == direction[keep].view('u4')).all(axis=1).mean():.4f}  "
          f"same pixels add a sample {(samples[..., 3] == pixels[..., 3]).mean():.5f}")
    # also behind mirror and glass chains: the specular BSDFs' directions carry the reference's bits since the
    # continuation is sampled with wo = -ray.d and the device's sinf / cosf are the host libm's
    assert np.array_equal(hemi[keep, 1:4].view("u4"), pos[keep].view("u4")), "hemi camera positions differ"
    assert np.array_equal(hemi[keep, 4:7].view("u4"), direction[keep].view("u4")), "hemi camera axes differ"
    assert np.array_equal(samples[..., 3], pixels[..., 3]), "which pixels add a sample differs"
    took = (pixels[..., 3] > 0) & (samples[..., 3] > 0)
    assert took.mean() > 0.5
    got, want = samples[..., :3][took], pixels[..., :3][took]
    rel = (np.abs(got - want) / np.maximum(np.abs(want), 1e-3)).max(axis=-1)
    frac = float((rel < 1e-4).mean())
    mean_ratio = float(got.mean() / want.mean())
    print(f"{variant} {rect}: {int(took.sum())} samples, {100 * frac:.2f} % within 1e-4, {100 * float((rel < 1e-2).mean()):.2f} % "
          f"within 1e-2, mean ratio {mean_ratio:.5f}, reference gather {info['gather_seconds']:.3f} s")
    assert want.mean() > 0
    assert frac >= 0.995, frac  # measured: 100 % on all four cases
    assert abs(mean_ratio - 1) < 1e-4, mean_ratio

    # ---- B: the whole chain against the reference's own RenderView + identity network --------------------
    hemi_b, films_b, pixels_b, _ = oracle.ref_iispt_task(flat, workdir, x0, y0, x1, y1, ts, SIZE, 1000, 3, 424242)
    assert films_b.shape == films_np.shape
    assert abs(float(films_np.mean()) / float(films_b.mean()) - 1) < 0.02
    assert abs(float(samples[..., :3].mean()) / float(pixels_b[..., :3].mean()) - 1) < 0.03


def test_iispt_task_with_a_torch_network_in_process(gpu, ref_scenes):
    """The hand-off SURVEY §8f.3 asks for: the [n,7,32,32] tensor the network reads and the [n,3,32,32] tensor it
    answers with never leave HBM — `net` is a torch module called on the tensor the hemisphere batch wrote
    (random weights: ml/iispt_net.py's IISPTNet has this contract; its trained weights are not in the tree)."""
    import torch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(7, 16, 3, padding=1), torch.nn.LeakyReLU(0.2),
                              torch.nn.Conv2d(16, 3, 1), torch.nn.ReLU()).cuda().eval()
    seen = {}

    def spy(x):
        seen["in"] = (tuple(x.shape), x.dtype, x.device.type, bool(torch.isfinite(x).all()))
        y = net(x)
        seen["out"] = tuple(y.shape)
        return y

    scene = gpu.Scene.from_pbrt(ref_scenes / "cfg1.pbrt")
    task = gpu.IisptTask(100, 100, 228, 228, 32, SIZE, 0, 0, 1)
    samples, pos, direction, valid, films = scene.iispt_task(task, spy)
    n = int(valid.sum())
    assert seen["in"] == ((n, 7, SIZE, SIZE), torch.float32, "cuda", True) and seen["out"] == (n, 3, SIZE, SIZE)
    assert samples.shape == (128, 128, 4) and np.isfinite(samples).all()
    assert (samples[..., 3] == 0.5).mean() > 0.5 and samples[..., :3].max() > 0


def test_iispt_task_rejects_bad_tasks(gpu, ref_scenes):
    scene = gpu.Scene.from_pbrt(ref_scenes / "cfg1.pbrt")
    for bad in (gpu.IisptTask(10, 10, 10, 20, 4, SIZE, 0, 0, 0), gpu.IisptTask(0, 0, 8, 8, 0, SIZE, 0, 0, 0),
                gpu.IisptTask(0, 0, 8, 8, 4, 0, 0, 0, 0)):
        with pytest.raises(gpu.IileError):
            scene.iispt_hemi_points(bad)
