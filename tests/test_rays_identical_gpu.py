"""Every ray of a render, bit for bit (needs a B200 and oracle/_ref on the box).

The device dumps every ray it traces during a render (`iile_render_dump_rays_dev`: camera, bounce, MIS and shadow
rays); the UNMODIFIED reference renders the same flattened scene with the same per-(pixel, sample) PCG32 streams
through an aggregate that logs each ray before forwarding it to its `BVHAccel` (`ref_harness render`,
`REF_LOG_RAYS`).  The two are compared as SETS of bit patterns of (origin, direction) — tools/compare_rays.py.
Identity of the sets means that every hit point, shading frame, BSDF sample, light sample and ray offset on the way
to every ray carried the reference's bits: `Triangle` / `Sphere::Intersect` and their shading geometry, `BSDF::Sample_f`
of every lobe, `Light::Sample_Li`, `SpawnRay[To]`, and the host libm the reference calls on the way (sinf / cosf /
acosf / atan2f, and the double sincos of `TrowbridgeReitzSample11`'s normal-incidence branch, microfacet.cpp:241-247)
— csrc/cuda/libm_exact.cuh.  Bar: bit-exact, not one ray different in either direction."""
import sys

import pytest

from conftest import ROOT
from test_render_gpu import gpu  # noqa: F401  (fixture)
from test_variants_gpu import variant

pytestmark = pytest.mark.gpu

sys.path.insert(0, str(ROOT / "tools"))

VARIANTS = ["substrate_wall", "glass_rough", "uber_full", "translucent_sheet", "oren_nayar_rough", "plastic_noremap",
            "sphere_light", "disk_light", "cylinder_instead_of_sphere", "spot_light", "infinite_light", "strategy_power"]


def check(out):
    assert out["device_rays"] == out["reference_rays"], out
    assert out["reference_rays_not_traced_by_the_device"] == 0, out
    for kind in ("camera", "bounce", "mis", "shadow"):
        if kind in out:
            assert out[kind]["not_traced_by_the_reference"] == 0, out
    assert out["camera"]["rays"] > 0 and out["bounce"]["rays"] > 0 and out["shadow"]["rays"] > 0, out


@pytest.mark.parametrize("name", VARIANTS)
def test_every_ray_of_a_variant_is_one_the_reference_traces(gpu, oracle, workdir, name):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    from compare_rays import compare
    src = workdir / f"rays_{name}.pbrt"
    src.write_text(variant(name))
    check(compare(src, depth=3, spp=4))


def test_every_ray_of_cfg1_is_one_the_reference_traces(gpu, oracle):
    """BASELINE configs[0] (killeroo-simple, 400x400) at 2 spp, maxdepth 3: 1.8 M rays."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    from compare_rays import compare
    check(compare(oracle.REF_SCENES / "cfg1.pbrt", depth=3, spp=2))
