"""IILE's direct-illumination pass: IisptRenderRunner::run_direct (iisptrenderrunner.cpp:601-631) runs
DirectProgressiveIntegrator::RenderOnePass (directprogressiveintegrator.cpp:65-150) iileDirectSamples times — one camera
sample per pixel and pass, Li = DirectLightingIntegrator's with UniformSampleAllLights and maxDepth 5 — into an
IisptFilmMonitor, which keeps a plain per-pixel sum of RGB and of ray weights (no pixel filter).

On the device that is `Integrator "directlighting" "integer maxdepth" [5]` with `pixelsamples` = the number of passes
and the film in IILE_FILTER_MONITOR mode: every sample is added with weight 1 to the pixel it was drawn for (a box
filter of half-width 0.5 does the same except for the one sample in ~10^5 whose `pixel + u` rounds up to the next
pixel in float — the two pixels of 6,144 that used to differ in this test).  Oracle: the UNMODIFIED
reference's DirectProgressiveIntegrator + IisptFilmMonitor (ref_harness direct_pass), pass p drawing the streams of
sample p.  Bar (floating point): >= 99.5 % of the pixels within 1e-5 relative, relMSE <= 1e-5 (test_render_gpu.py's)."""
import numpy as np
import pytest

from conftest import ROOT
from test_render_gpu import film_agreement, gpu  # noqa: F401  (gpu is a fixture)

pytestmark = pytest.mark.gpu

SMOKE = (ROOT / "scenes" / "smoke.pbrt").read_text()


@pytest.mark.parametrize("name,passes,edits", [
    ("plain", 6, ()),
    ("samples3_uber", 4, (('AreaLightSource "diffuse" "color L" [12 11 9]', 'AreaLightSource "diffuse" "color L" [12 11 9] "integer samples" [3]'),
                          ('Material "glass" "float eta" [1.5]', 'Material "uber" "color Kr" [.3 .3 .3] "color Kt" [.5 .5 .6] "color Kd" [.1 .1 .1]'))),
    ("interface_crop", 3, (('Material "glass" "float eta" [1.5]', 'Material "none"'),
                           ('"string filename" "smoke.pfm"', '"string filename" "smoke.pfm" "float cropwindow" [.1 .8 .2 .9]'))),
])
def test_direct_pass_matches_directprogressive_and_film_monitor(gpu, oracle, workdir, name, passes, edits):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    text = SMOKE.replace('Integrator "path" "integer maxdepth" [6]', 'Integrator "directlighting" "integer maxdepth" [5]')
    text = text.replace('Sampler "random" "integer pixelsamples" [8]',
                        f'Sampler "random" "integer pixelsamples" [{passes}]\nPixelFilter "box" "float xwidth" [.5] "float ywidth" [.5]')
    for old, new in edits:
        assert old in text
        text = text.replace(old, new)
    src = workdir / f"direct_pass_{name}.pbrt"
    src.write_text(text)
    scene = gpu.Scene.from_pbrt(src)
    flat = workdir / f"direct_pass_{name}.flat"
    scene.save_flat(flat)
    setup = scene.info().setup
    setup.filterType = gpu.FILTER_MONITOR  # a sample lands on the pixel it was drawn for (IisptFilmMonitor::add_n_samples)
    scene.set_setup(setup)
    rgb = scene.resolve(scene.render()[0])
    ref, info = oracle.ref_direct_pass(flat, workdir, passes)
    assert ref.shape == rgb.shape, (ref.shape, rgb.shape)
    frac, relmse = film_agreement(rgb, ref)
    print(f"{name}: {passes} passes, {100 * frac:.2f} % of pixels within 1e-5, relMSE {relmse:.3g}, "
          f"mean ratio {float(rgb.mean() / ref.mean()):.7f}, reference {info['seconds']:.2f} s")
    assert ref.mean() > 0.05 and frac >= 0.995 and relmse <= 1e-5, (frac, relmse)
