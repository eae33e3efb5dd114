"""Parameter / edge-case parity sweep (needs a B200 and oracle/_ref on the box).

Each case is scenes/smoke.pbrt with one directive changed — the parameters BASELINE's north_star
names (`maxdepth`, `rrthreshold`, `pixelsamples`, crop window) and the plugin parameters either
side of them (camera lens, screen window, every pixel filter, light-sampling strategies, BVH
split methods / leaf sizes, partial spheres, two-sided and reversed emitters, Oren-Nayar sigma,
rough plastic without roughness remapping).  The same file text goes to both sides: the GPU path
parses it itself; the reference side is the UNMODIFIED reference run on the flattened scene
(`ref_harness render`) with the same per-(pixel, sample) PCG32 streams, or — for the halton
cases — with its own stock HaltonSampler.  Bars are test_render_gpu.py's: >= 99.5 % of pixels
within 1e-5 relative, relMSE <= 1e-5, ray counters within 0.2 %.
"""
import numpy as np
import pytest

from conftest import ROOT
from test_render_gpu import check_counts, film_agreement, gpu  # noqa: F401  (gpu is a fixture)

pytestmark = pytest.mark.gpu

SMOKE = (ROOT / "scenes" / "smoke.pbrt").read_text()
CAMERA = 'Camera "perspective" "float fov" [50]'
SAMPLER = 'Sampler "random" "integer pixelsamples" [8]'
INTEGRATOR = 'Integrator "path" "integer maxdepth" [6]'
ACCEL = 'Accelerator "bvh" "integer maxnodeprims" [4]'
FILM = 'Film "image" "integer xresolution" [96] "integer yresolution" [64] "string filename" "smoke.pfm"'
SPHERE = 'Shape "sphere" "float radius" [.45]'
AREA = 'AreaLightSource "diffuse" "color L" [12 11 9]'
PLASTIC = 'Material "plastic" "color Kd" [.3 .3 .7] "color Ks" [.4 .4 .4] "float roughness" [.08]'
MATTE_SIGMA = '"float sigma" [20]'
GLASS = 'Material "glass" "float eta" [1.5]'
MIRROR = 'Material "mirror" "color Kr" [.9 .9 .9]'
POINT = 'LightSource "point" "color I" [3 3 4] "point from" [-1.8 -2.5 2.2]'
CEILING = '"point P" [-2.5 -5 3  2.5 -5 3  2.5 2.5 3  -2.5 2.5 3]'
OPEN_CEILING = '"point P" [-.6 -1 3  .6 -1 3  .6 .4 3  -.6 .4 3]'
RED_WALL = 'Material "matte" "color Kd" [.7 .2 .2] "float sigma" [20]'

CASES = {
    # Integrator "path" parameters
    "maxdepth0": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [0]'),
    "maxdepth1": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [1]'),
    "maxdepth12_rr": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [12] "float rrthreshold" [0.25]'),
    "rrthreshold_high": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [8] "float rrthreshold" [50]'),
    "pixelbounds": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [4] "integer pixelbounds" [7 61 3 40]'),
    "strategy_uniform": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "uniform"'),
    "strategy_power": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "power"'),
    # Integrator "directlighting" (directlighting.cpp; the Li of IILE's DirectProgressiveIntegrator): all / one
    # light per vertex, the specular recursion through the mirror and the glass prism walked depth-first
    "direct_all": (INTEGRATOR, 'Integrator "directlighting"'),
    "direct_one_depth3": (INTEGRATOR, 'Integrator "directlighting" "string strategy" "one" "integer maxdepth" [3]'),
    "direct_all_depth1_pixelbounds": (INTEGRATOR, 'Integrator "directlighting" "integer maxdepth" [1] "integer pixelbounds" [5 80 2 50]'),
    "direct_all_depth8_uber": (INTEGRATOR, 'Integrator "directlighting" "integer maxdepth" [8]',
                               (GLASS, 'Material "uber" "color Kd" [.1 .1 .1] "color Ks" [.2 .2 .2] "color Kr" [.4 .4 .4] '
                                       '"color Kt" [.5 .5 .5] "float index" [1.3]')),
    "direct_one_infinite": (INTEGRATOR, 'Integrator "directlighting" "string strategy" "one"',
                            (POINT, 'LightSource "infinite" "color L" [.5 .6 .9]'), (CEILING, OPEN_CEILING)),
    # sample arrays (integrator.cpp:54-83): "samples" > 1 on the area light => that many EstimateDirect terms per
    # vertex from the sampler's 2D arrays while the path has arrays left (maxdepth levels), one from Get2D after
    "direct_all_samples4": (INTEGRATOR, 'Integrator "directlighting"', (AREA, AREA + ' "integer samples" [4]')),
    "direct_all_nsamples3_depth2_sphere_light": (INTEGRATOR, 'Integrator "directlighting" "integer maxdepth" [2]',
                                                 (AREA, AREA + ' "integer nsamples" [3]'),
                                                 (SPHERE, 'AreaLightSource "diffuse" "color L" [6 5 4] "integer samples" [2]\n  Shape "sphere" "float radius" [.3]')),
    "path_ignores_samples": (INTEGRATOR, 'Integrator "path" "integer maxdepth" [4]', (AREA, AREA + ' "integer samples" [4]')),
    # the direct integrators under the GlobalSamplers (sampler.cpp:128-196): the arrays live in dimensions 5 .. and
    # the rest of the vertex's numbers continue behind them — compared with the STOCK reference samplers
    "direct_all_halton": (INTEGRATOR, 'Integrator "directlighting"', (SAMPLER, 'Sampler "halton" "integer pixelsamples" [8]')),
    "direct_all_halton_samples4": (INTEGRATOR, 'Integrator "directlighting" "integer maxdepth" [3]',
                                   (SAMPLER, 'Sampler "halton" "integer pixelsamples" [4]'), (AREA, AREA + ' "integer samples" [4]')),
    "direct_one_halton": (INTEGRATOR, 'Integrator "directlighting" "string strategy" "one"',
                          (SAMPLER, 'Sampler "halton" "integer pixelsamples" [8]')),
    "direct_all_sobol_samples3": (INTEGRATOR, 'Integrator "directlighting" "integer maxdepth" [4]',
                                  (SAMPLER, 'Sampler "sobol" "integer pixelsamples" [8]'), (AREA, AREA + ' "integer samples" [3]')),
    "direct_all_sobol": (INTEGRATOR, 'Integrator "directlighting" "integer maxdepth" [4]',
                         (SAMPLER, 'Sampler "sobol" "integer pixelsamples" [8]')),
    "direct_one_sobol": (INTEGRATOR, 'Integrator "directlighting" "string strategy" "one"',
                         (SAMPLER, 'Sampler "sobol" "integer pixelsamples" [4]')),
    "whitted_halton": (INTEGRATOR, 'Integrator "whitted"', (SAMPLER, 'Sampler "halton" "integer pixelsamples" [8]')),
    # Integrator "whitted" (whitted.cpp:44-88): unweighted loop over the lights + the same specular recursion
    "whitted": (INTEGRATOR, 'Integrator "whitted"'),
    "whitted_depth2_spot_distant": (INTEGRATOR, 'Integrator "whitted" "integer maxdepth" [2]',
                                    (POINT, 'LightSource "spot" "color I" [30 30 40] "point from" [-1.8 -2.5 2.2] "point to" [0.4 0 0.3]\n'
                                            'LightSource "distant" "color L" [.4 .4 .4] "point from" [1 -3 4] "point to" [0 0 0]')),
    # Sampler / Film / PixelFilter
    "spp1": (SAMPLER, 'Sampler "random" "integer pixelsamples" [1]'),
    "spp33": (SAMPLER, 'Sampler "random" "integer pixelsamples" [33]'),
    "cropwindow": (FILM, FILM + ' "float cropwindow" [.2 .77 .31 .9]'),
    "odd_resolution": (FILM, 'Film "image" "integer xresolution" [101] "integer yresolution" [47] "string filename" "o.pfm"'),
    "film_scale_maxlum": (FILM, FILM + ' "float scale" [2.5] "float maxsampleluminance" [1.5]'),
    "filter_gaussian": (SAMPLER, SAMPLER + '\nPixelFilter "gaussian" "float xwidth" [1.5] "float ywidth" [2.5] "float alpha" [1.2]'),
    "filter_mitchell": (SAMPLER, SAMPLER + '\nPixelFilter "mitchell"'),
    "filter_triangle": (SAMPLER, SAMPLER + '\nPixelFilter "triangle" "float xwidth" [1.7]'),
    "filter_sinc": (SAMPLER, SAMPLER + '\nPixelFilter "sinc" "float tau" [2]'),
    "filter_box_wide": (SAMPLER, SAMPLER + '\nPixelFilter "box" "float xwidth" [1.25] "float ywidth" [.75]'),
    "halton_crop": (SAMPLER, 'Sampler "halton" "integer pixelsamples" [5]',
                    (FILM, FILM + ' "float cropwindow" [.1 .6 .25 1]')),
    "halton_pixelcenter_mitchell": (SAMPLER, 'Sampler "halton" "integer pixelsamples" [4] "bool samplepixelcenter" "true"\n'
                                             'PixelFilter "mitchell"'),
    "halton_lens": (SAMPLER, 'Sampler "halton" "integer pixelsamples" [6]',
                    (CAMERA, 'Camera "perspective" "float fov" [50] "float lensradius" [.08] "float focaldistance" [4.5]')),
    # Camera "perspective"
    "lens": (CAMERA, 'Camera "perspective" "float fov" [50] "float lensradius" [.06] "float focaldistance" [4]'),
    "halffov_aspect": (CAMERA, 'Camera "perspective" "float halffov" [31] "float frameaspectratio" [1.2]'),
    "screenwindow": (CAMERA, 'Camera "perspective" "float fov" [55] "float screenwindow" [-1.2 .9 -.8 .7]'),
    "camera_orthographic": (CAMERA, 'Camera "orthographic" "float screenwindow" [-2.6 2.6 -1.8 1.8]'),
    "camera_orthographic_lens": (CAMERA, 'Camera "orthographic" "float frameaspectratio" [1.4] "float lensradius" [.05] "float focaldistance" [4]'),
    "camera_environment": (CAMERA, 'Camera "environment"', (FILM, 'Film "image" "integer xresolution" [128] "integer yresolution" [64] "string filename" "e.pfm"')),
    # Accelerator "bvh"
    "bvh_leaf1": (ACCEL, 'Accelerator "bvh" "integer maxnodeprims" [1]'),
    "bvh_leaf16": (ACCEL, 'Accelerator "bvh" "integer maxnodeprims" [16]'),
    "bvh_middle": (ACCEL, 'Accelerator "bvh" "string splitmethod" "middle"'),
    "bvh_hlbvh": (ACCEL, 'Accelerator "bvh" "string splitmethod" "hlbvh"'),
    "bvh_hlbvh_leaf2": (ACCEL, 'Accelerator "bvh" "string splitmethod" "hlbvh" "integer maxnodeprims" [2]'),
    "bvh_equal": (ACCEL, 'Accelerator "bvh" "string splitmethod" "equal" "integer maxnodeprims" [2]'),
    # shapes, lights, materials
    # Shape "disk" (disk.cpp): as an occluder, as a partial annulus, and as the area light (Disk::Sample)
    "disk_instead_of_sphere": (SPHERE, 'Rotate 40 1 0 0\n  Shape "disk" "float radius" [.6] "float height" [.1]'),
    "disk_annulus_partial": (SPHERE, 'Rotate -25 0 1 0\n  Shape "disk" "float radius" [.7] "float innerradius" [.25] "float phimax" [200]'),
    "disk_light": (SPHERE, 'AreaLightSource "diffuse" "color L" [8 7 6]\n  Rotate 150 1 0 0\n  Shape "disk" "float radius" [.35]'),
    "disk_light_reversed_twosided": (SPHERE, 'AreaLightSource "diffuse" "color L" [4 4 6] "bool twosided" "true"\n  ReverseOrientation\n'
                                             '  Rotate 70 0 1 0\n  Shape "disk" "float radius" [.3] "float height" [.2]'),
    "cylinder_instead_of_sphere": (SPHERE, 'Rotate 55 1 0 0\n  Shape "cylinder" "float radius" [.3] "float zmin" [-.4] "float zmax" [.5]'),
    "cylinder_light_partial": (SPHERE, 'AreaLightSource "diffuse" "color L" [6 6 5]\n  Rotate 80 0 1 0\n'
                                       '  Shape "cylinder" "float radius" [.12] "float zmin" [-.5] "float zmax" [.5] "float phimax" [270]'),
    "partial_sphere": (SPHERE, 'Shape "sphere" "float radius" [.45] "float zmin" [-.2] "float zmax" [.33] "float phimax" [250]'),
    "sphere_reversed": (SPHERE, 'ReverseOrientation\n  Shape "sphere" "float radius" [.45]'),
    "sphere_light": (SPHERE, 'AreaLightSource "diffuse" "color L" [6 5 4]\n  Shape "sphere" "float radius" [.3]'),
    "twosided_light": (AREA, 'AreaLightSource "diffuse" "color L" [12 11 9] "bool twosided" "true" "float scale" [.5]\n  Translate 0 0 -.5'),
    "area_alias": (AREA, 'AreaLightSource "area" "color L" [3 9 12]'),
    "plastic_noremap": (PLASTIC, 'Material "plastic" "color Kd" [.3 .3 .7] "color Ks" [.4 .4 .4] "float roughness" [.2] '
                                 '"bool remaproughness" "false"'),
    "lambertian_wall": (MATTE_SIGMA, '"float sigma" [0]'),
    "oren_nayar_rough": (MATTE_SIGMA, '"float sigma" [75]'),
    # rough glass (MicrofacetReflection + MicrofacetTransmission, glass.cpp:66-91) and uber (uber.cpp:45-104)
    "glass_rough": (GLASS, 'Material "glass" "float eta" [1.5] "float uroughness" [.2] "float vroughness" [.2]'),
    "glass_rough_aniso_noremap": (GLASS, 'Material "glass" "float index" [1.33] "float uroughness" [.25] "float vroughness" [.06] '
                                         '"bool remaproughness" "false" "color Kt" [.9 .95 1]'),
    "glass_rough_wall": (RED_WALL,
                         'Material "glass" "float eta" [1.4] "float uroughness" [.5] "float vroughness" [.5] "color Kr" [.6 .6 .6] "color Kt" [0 0 0]'),
    "glass_reflect_only": (GLASS, 'Material "glass" "color Kt" [0 0 0] "float uroughness" [.1] "float vroughness" [.1]'),
    # substrate (FresnelBlend, substrate.cpp) and translucent (diffuse + glossy reflection and transmission)
    "substrate_default": (PLASTIC, 'Material "substrate"'),
    "substrate_aniso_noremap": (PLASTIC, 'Material "substrate" "color Kd" [.3 .5 .2] "color Ks" [.1 .1 .1] "float uroughness" [.3] '
                                         '"float vroughness" [.05] "bool remaproughness" "false"'),
    "substrate_wall": (RED_WALL, 'Material "substrate" "color Kd" [.6 .2 .2] "color Ks" [.3 .3 .3] "float uroughness" [.02] "float vroughness" [.02]'),
    "translucent_default": (GLASS, 'Material "translucent"'),
    "translucent_sheet": (MIRROR, 'Material "translucent" "color Kd" [.6 .6 .3] "color Ks" [.3 .3 .3] "color reflect" [.2 .2 .2] '
                                  '"color transmit" [.8 .8 .8] "float roughness" [.3]'),
    "translucent_reflect_only": (GLASS, 'Material "translucent" "color transmit" [0 0 0] "color Kd" [.5 .5 .5]'),
    "uber_default": (PLASTIC, 'Material "uber"'),
    "uber_full": (GLASS, 'Material "uber" "color Kd" [.2 .3 .1] "color Ks" [.3 .3 .3] "color Kr" [.2 .2 .2] "color Kt" [.4 .4 .5] '
                         '"float roughness" [.05] "float index" [1.4] "color opacity" [.8 .8 .8]'),
    "uber_aniso_mirror": (MIRROR, 'Material "uber" "color Kd" [0 0 0] "color Ks" [.6 .6 .6] "color Kr" [.3 .3 .3] '
                                  '"float uroughness" [.3] "float vroughness" [.02] "bool remaproughness" "false"'),
    # delta lights besides "point" (spot.cpp, distant.cpp), under every light-sampling strategy
    "spot_light": (POINT, 'LightSource "spot" "color I" [30 30 40] "point from" [-1.8 -2.5 2.2] "point to" [0.4 0 0.3] '
                          '"float coneangle" [35] "float conedeltaangle" [12]'),
    "spot_light_rotated_power": (POINT, 'AttributeBegin\n  Rotate 20 0 1 0\n  Translate .2 0 .1\n  LightSource "spot" "color I" [25 25 25] '
                                        '"point from" [-1 -2 2.4] "point to" [0 0 0] "float coneangle" [50] "color scale" [1 .8 .6]\nAttributeEnd',
                                 (INTEGRATOR, 'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "power"')),
    "distant_light": (POINT, 'LightSource "distant" "color L" [1.5 1.4 1.2] "point from" [-1 -3 4] "point to" [0 0 0]'),
    "distant_light_uniform": (POINT, 'AttributeBegin\n  Rotate -15 1 0 0\n  LightSource "distant" "color L" [2 2 2] "point from" [0 -1 2]\nAttributeEnd',
                              (INTEGRATOR, 'Integrator "path" "integer maxdepth" [4] "string lightsamplestrategy" "uniform"')),
    # constant InfiniteAreaLight (infinite.cpp) seen through an opened ceiling: escaped camera / specular
    # rays, light sampling through the 2x2 Distribution2D, BSDF-sampled rays that escape (MIS)
    "infinite_light": (POINT, 'LightSource "infinite" "color L" [.5 .6 .9]', (CEILING, OPEN_CEILING)),
    "infinite_light_rotated_only": (POINT, 'AttributeBegin\n  Rotate 35 1 0 0\n  Rotate 20 0 0 1\n  LightSource "infinite" "color L" [1 1 1] '
                                           '"color scale" [.8 .7 .5]\nAttributeEnd',
                                    (CEILING, OPEN_CEILING), (AREA, 'AreaLightSource "diffuse" "color L" [0 0 0]')),
    "infinite_light_power": (POINT, 'LightSource "infinite" "color L" [.3 .3 .3]', (CEILING, OPEN_CEILING),
                             (INTEGRATOR, 'Integrator "path" "integer maxdepth" [5] "string lightsamplestrategy" "power"')),
    # interface-only primitives (Material "none"): a crossing is a step of the path that counts no bounce
    # (path.cpp:111-116, directlighting.cpp:74-75) — the paths that cross need more wavefront steps than maxdepth + 1 —
    # and shadow / MIS rays are stopped by it like by any other surface
    "interface_prism": (GLASS, 'Material "none"'),
    "interface_prism_depth1": (GLASS, 'Material "none"', (INTEGRATOR, 'Integrator "path" "integer maxdepth" [1]')),
    "interface_prism_direct": (GLASS, 'Material "none"', (INTEGRATOR, 'Integrator "directlighting" "integer maxdepth" [3]')),
    "interface_sphere_halton": ('Material "matte" "color Kd" [.8 .7 .3]', 'Material "none"',
                                (SAMPLER, 'Sampler "halton" "integer pixelsamples" [8]')),
    # a scene without lights renders black (and must not divide by the light count)
    "no_lights": (POINT, '', (AREA, '')),
    "uber_transparent": (GLASS, 'Material "uber" "color Kd" [.5 .5 .5] "color opacity" [0 0 0]'),
}


def variant(name):
    text = SMOKE
    spec = CASES[name]
    for old, new in [(spec[0], spec[1])] + list(spec[2:]):
        assert old in text, (name, old)
        text = text.replace(old, new, 1)
    return text


@pytest.mark.parametrize("name", sorted(CASES))
def test_variant_matches_the_reference(gpu, oracle, workdir, name):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    src = workdir / f"variant_{name}.pbrt"
    src.write_text(variant(name))
    scene = gpu.Scene.from_pbrt(src)
    flat = workdir / f"variant_{name}.flat"
    scene.save_flat(flat)
    film, stats = scene.render()
    rgb = scene.resolve(film)
    halton = scene.info().setup.samplerType in (gpu.SAMPLER_HALTON, gpu.SAMPLER_SOBOL)   # the stock GlobalSamplers
    out = workdir / f"variant_{name}.pfm"
    oracle.ref_render(flat, out, sampler="stock" if halton else "pixelstream")
    ref = oracle.read_pfm(out)
    assert ref.shape == rgb.shape, (ref.shape, rgb.shape)
    assert np.isfinite(rgb).all()
    frac, relmse = film_agreement(rgb, ref)
    assert frac >= 0.995 and relmse <= 1e-5, (name, frac, relmse)
    st = oracle.ref_stats(out)
    check_counts(stats, np.array([st["cameraRays"], st["closestRays"], st["shadowRays"]]))


def test_deep_maxdepth_ends_early_without_changing_the_film(gpu, workdir):
    """maxdepth 100: Russian roulette has ended every path long before step 100; the batch stops launching once its
    queue is empty (a read-back every fourth step from depth 8 on).  Same film as maxdepth 40, far fewer launches than
    the 900 a hundred steps would take."""
    films, launches = [], []
    for depth in (40, 100):
        src = workdir / f"deep_{depth}.pbrt"
        src.write_text(SMOKE.replace(INTEGRATOR, f'Integrator "path" "integer maxdepth" [{depth}]'))
        scene = gpu.Scene.from_pbrt(src)
        film, stats = scene.render()
        films.append(film)
        launches.append(stats.launches)
    assert launches[1] < 450, launches
    # both runs end by roulette well before step 40, so the films agree (a path alive at step 40 would differ)
    assert np.allclose(films[0], films[1], rtol=1e-6, atol=1e-7)
