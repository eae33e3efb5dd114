"""The whole IILE render — IISPTIntegrator::render_normal_2 (iispt.cpp:357-447) — through Scene.iile_render:
IisptScheduleMonitor's tasks, each one hemi points -> RenderView -> network (in process) -> gather into the indirect film
monitor; DirectProgressiveIntegrator's passes into the direct one; merge_into.

Oracle: `ref_harness iile_render` — the UNMODIFIED reference's IisptScheduleMonitor, IisptRenderRunner members,
IISPTdIntegrator, DirectProgressiveIntegrator and IisptFilmMonitor chained on one thread with the identity network and
the streams the device draws from (oracle/ref_harness/iispt_task.cpp).  The schedule itself (pure integer logic) is
compared task by task on the CPU.  Bars (floating point, the chain includes RenderView's sinf / cosf and the
network-side log / exp): the direct film >= 97 % of the pixels within 1e-5; the merged film >= 95 % of the pixels
within 1e-3 relative and its mean within 0.5 %."""
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


def test_iile_render_matches_the_reference_chain(gpu, oracle, workdir):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    scene = gpu.Scene.from_flat(ROOT / "tests" / "golden" / "smoke.flat")  # 96 x 64, default box filter
    tasks, passes, size, radius, seed = 6, 4, 16, 6.0, 99
    merged, direct, indirect = scene.iile_render(lambda x: x[:, 0:3], indirect_tasks=tasks, direct_samples=passes,
                                                 hemi_size=size, radius_start=radius, seed=seed)
    ref, info = oracle.ref_iile_render(ROOT / "tests" / "golden" / "smoke.flat", workdir, tasks, passes, size, radius, seed)
    assert info["tasks"] == tasks and ref.shape == merged.shape == (64, 96, 3)
    ref_direct, _ = oracle.ref_direct_pass(ROOT / "tests" / "golden" / "smoke.flat", workdir, passes)
    rel_d = (np.abs(direct - ref_direct) / np.maximum(np.abs(ref_direct), 1e-2)).max(axis=-1)
    rel = (np.abs(merged - ref) / np.maximum(np.abs(ref), 1e-2)).max(axis=-1)
    print(f"direct: {100 * float((rel_d < 1e-5).mean()):.2f} % of pixels within 1e-5; merged: "
          f"{100 * float((rel < 1e-5).mean()):.2f} % within 1e-5, {100 * float((rel < 1e-3).mean()):.2f} % within 1e-3, "
          f"mean ratio {float(merged.mean() / ref.mean()):.6f}; reference chain {info['indirect_seconds']:.1f} s")
    assert (rel_d < 1e-5).mean() >= 0.97
    assert (rel < 1e-3).mean() >= 0.95 and abs(float(merged.mean() / ref.mean()) - 1) < 5e-3
    assert indirect.mean() > 0.01 and direct.mean() > 0.05  # both components are there
    # a second sweep of the schedule happened (6 tasks > the 4 of one sweep): the weight of a pixel is 0.5 per sample,
    # so its mean is twice the samples' (add_n_samples adds the RGB unweighted, iisptfilmmonitor.cpp:47-75)
    assert np.isfinite(merged).all()
