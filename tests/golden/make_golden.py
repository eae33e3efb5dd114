"""Regenerates tests/golden/* from the REFERENCE (run where /root/reference and oracle/_ref
exist: `python tests/golden/make_golden.py`).  Inputs are seeded; outputs are small.

  pcg32.npz       RNG::UniformFloat streams from the reference's src/core/rng.h (via ref_harness genrays'
                  RNG, cross-checked here against the restatement)
  smoke.flat      scenes/smoke.pbrt flattened by the product front-end
  cast_smoke.npz  8192 rays + BVHAccel::Intersect / IntersectP answers of the reference on that scene,
                  the reference's LinearBVHNode array and primitive order, and the restatement's
                  node/test counts
  render_ref.npz  films of the reference's PathIntegrator (ref_harness render, "pixelstream" sampler: one
                  PCG32 stream per (pixel, sample), the streams the GPU path draws) on scenes/smoke.pbrt at
                  16 spp and on the three analytic furnace scenes, plus the reference's ray counters
  hemi_ref.npz    IISPTdIntegrator::RenderView of the reference (ref_harness hemi) for 8 hemispheric cameras
                  on scenes/smoke.pbrt: 32x32 intensity, camera-space normals, distance
  halton_ref.npz  film of the reference's PathIntegrator with its own, unmodified HaltonSampler ("stock" mode of
                  ref_harness render = what MakeSampler builds for `Sampler "halton"`) on scenes/smoke.pbrt at
                  8 spp, box filter and gaussian filter (negative sample-bounds origin), plus ray counters
                  (`python tests/golden/make_golden.py halton` regenerates only this file)
  sobol_ref.npz   the same two films under the reference's unmodified SobolSampler
                  (`python tests/golden/make_golden.py sobol` regenerates only this file)
  hemi_nn_ref.npz the reference's IisptRenderRunner::normalizeMapsDownstream / transformMapsUpstream around
                  RenderView (ref_harness hemi_nn) on the 8 hemispheres of hemi_ref.npz: the 7x32x32 network
                  input, the channel means, and the upstream transform of the identity network's answer
                  (`python tests/golden/make_golden.py hemi_nn` regenerates only this file)
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from conftest import load_product  # noqa: E402

GOLD = Path(__file__).resolve().parent


def halton(iile, wd):
    """Sampler "halton": the GPU draws the reference's own points, so the stock sampler is the oracle."""
    out = {}
    for tag, filt in (("box", None), ("gaussian", (iile.FILTER_GAUSSIAN, 2.0, 2.0, 2.0))):
        s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
        setup = s.info().setup
        setup.samplerType = iile.SAMPLER_HALTON
        setup.pixelSamples = 8
        if filt:
            setup.filterType = filt[0]
            setup.filterRadius[:] = [filt[1], filt[2]]
            setup.filterParam[0] = filt[3]
        s.set_setup(setup)
        s.save_flat(wd / f"halton_{tag}.flat")
        oracle.ref_render(wd / f"halton_{tag}.flat", wd / f"halton_{tag}.pfm", sampler="stock")
        out[tag + "_rgb"] = oracle.read_pfm(wd / f"halton_{tag}.pfm")
        st = oracle.ref_stats(wd / f"halton_{tag}.pfm")
        out[tag + "_counts"] = np.array([st["cameraRays"], st["closestRays"], st["shadowRays"]], dtype=np.int64)
    np.savez_compressed(GOLD / "halton_ref.npz", **out)


def sobol(iile, wd):
    """Sampler "sobol" (IILE's IILE_PATH_SAMPLES_OVERRIDE sampler): also a GlobalSampler, so the stock
    SobolSampler of the reference is the oracle, as for halton."""
    out = {}
    for tag, filt in (("box", None), ("gaussian", (iile.FILTER_GAUSSIAN, 2.0, 2.0, 2.0))):
        s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
        setup = s.info().setup
        setup.samplerType = iile.SAMPLER_SOBOL
        setup.pixelSamples = 8
        if filt:
            setup.filterType = filt[0]
            setup.filterRadius[:] = [filt[1], filt[2]]
            setup.filterParam[0] = filt[3]
        s.set_setup(setup)
        s.save_flat(wd / f"sobol_{tag}.flat")
        oracle.ref_render(wd / f"sobol_{tag}.flat", wd / f"sobol_{tag}.pfm", sampler="stock")
        out[tag + "_rgb"] = oracle.read_pfm(wd / f"sobol_{tag}.pfm")
        st = oracle.ref_stats(wd / f"sobol_{tag}.pfm")
        out[tag + "_counts"] = np.array([st["cameraRays"], st["closestRays"], st["shadowRays"]], dtype=np.int64)
    np.savez_compressed(GOLD / "sobol_ref.npz", **out)


def hemi_nn(wd):
    """hemi_nn_ref.npz: the reference's normalizeMapsDownstream / transformMapsUpstream
    (IisptRenderRunner, iisptrenderrunner.cpp:1041-1136) on the hemispheres of hemi_ref.npz."""
    g = np.load(GOLD / "hemi_ref.npz")
    nn, means, up = oracle.ref_hemispheres_nn(GOLD / "smoke.flat", g["pos"], g["dir"], wd)
    np.savez_compressed(GOLD / "hemi_nn_ref.npz", nn_input=nn, means=means, upstream=up)


def main():
    if sys.argv[1:] == ["halton"]:
        halton(load_product(), Path(tempfile.mkdtemp()))
        return
    if sys.argv[1:] == ["sobol"]:
        sobol(load_product(), Path(tempfile.mkdtemp()))
        return
    if sys.argv[1:] == ["hemi_nn"]:
        hemi_nn(Path(tempfile.mkdtemp()))
        return
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    iile = load_product()
    wd = Path(tempfile.mkdtemp())
    s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
    s.save_flat(GOLD / "smoke.flat")
    flat = oracle.FlatScene(GOLD / "smoke.flat")
    lo, hi = oracle.cfg2_ray_box(flat)
    # rays from the reference's own RNG + UniformSampleSphere (ref_harness genrays)
    oracle._run([oracle.REF_HARNESS, "genrays", GOLD / "smoke.flat", 8192, wd / "rays.bin"])
    rays = np.fromfile(wd / "rays.bin", dtype=oracle.RAY_DTYPE)
    mine = oracle.generate_rays(lo, hi, 0, 8192)
    assert rays.tobytes() == mine.tobytes(), "restated ray generator differs from the reference's RNG/sampling"
    # half the rays become finite shadow-style segments
    rays["tMax"][1::2] = np.float32(3.0)
    ref, _ = oracle.ref_raycast(GOLD / "smoke.flat", rays, wd)
    occ, _ = oracle.ref_raycast(GOLD / "smoke.flat", rays, wd, any_hit=True)
    nodes, order, _ = oracle.ref_bvh(GOLD / "smoke.flat", wd)
    _, nf, pt = oracle.cast(flat, rays)
    np.savez_compressed(GOLD / "cast_smoke.npz", rays=rays.view("<f4").reshape(-1, 8), prim=ref["prim"],
                        t=ref["t"], b1=ref["b1"], b2=ref["b2"], p=ref["p"], n=ref["n"], occluded=occ,
                        ref_nodes=nodes, ref_order=order, nodes_fetched=nf, prim_tests=pt)
    # PCG32 streams: the genrays output above pins sequence 0..; store raw floats of a few sequences
    seqs = np.array([0, 1, 624, (5 << 32) | 12345], dtype=np.uint64)
    floats = np.stack([oracle.pcg32_floats(int(q), 0, 64) for q in seqs])
    # cross-check sequence 0 against the reference-generated rays: first 5 draws build ray 0
    u = floats[0][:5]
    o = (1 - u[:3]) * lo + u[:3] * hi
    assert np.array_equal(o.astype("f4").view("u4"), rays["o"][0].view("u4"))
    np.savez_compressed(GOLD / "pcg32.npz", seqs=seqs, floats=floats)
    # ---- render goldens: the reference's PathIntegrator with one PCG32 stream per (pixel, sample) ----
    rg = {}
    setup = s.info().setup
    setup.pixelSamples = 16
    s.set_setup(setup)
    s.save_flat(wd / "smoke16.flat")
    oracle.ref_render(wd / "smoke16.flat", wd / "smoke16.pfm", sampler="pixelstream")
    rg["smoke_rgb"] = oracle.read_pfm(wd / "smoke16.pfm")
    st = oracle.ref_stats(wd / "smoke16.pfm")
    rg["smoke_counts"] = np.array([st["cameraRays"], st["closestRays"], st["shadowRays"]], dtype=np.int64)
    for name in ("furnace_point", "furnace_4points", "furnace_area"):
        f = iile.Scene.from_pbrt(ROOT / "scenes" / f"{name}.pbrt", upload=False)
        f.save_flat(wd / f"{name}.flat")
        oracle.ref_render(wd / f"{name}.flat", wd / f"{name}.pfm", sampler="pixelstream")
        rg[name + "_rgb"] = oracle.read_pfm(wd / f"{name}.pfm")
        st = oracle.ref_stats(wd / f"{name}.pfm")
        rg[name + "_counts"] = np.array([st["cameraRays"], st["closestRays"], st["shadowRays"]], dtype=np.int64)
    np.savez_compressed(GOLD / "render_ref.npz", **rg)
    # ---- hemispheric auxiliary pass (BASELINE configs[3]): 8 hemispheres at hit points of the golden rays ----
    g = np.load(GOLD / "cast_smoke.npz")
    hit = np.flatnonzero((g["prim"] >= 0) & np.isinf(g["rays"][:, 3]))[:8]
    n = g["n"][hit].astype(np.float32)
    d = g["rays"][hit, 4:7]
    n = np.where((n * d).sum(axis=1, keepdims=True) > 0, -n, n)          # face the incoming ray
    pos = (g["p"][hit] + np.float32(0.01) * n).astype(np.float32)
    inten, nrm, dist, _ = oracle.ref_hemispheres(GOLD / "smoke.flat", pos, n, wd)
    np.savez_compressed(GOLD / "hemi_ref.npz", pos=pos, dir=n, intensity=inten, normals=nrm, distance=dist)
    hemi_nn(wd)
    halton(iile, wd)
    sobol(iile, wd)
    print("golden written:", [p.name for p in GOLD.iterdir()])


if __name__ == "__main__":
    main()
