"""Image textures, bump mapping and ray differentials on the device (csrc/cuda/texture.cuh; SURVEY §8 a14 + f2,
VERDICT r1 missing #1 / #2) against the STOCK reference binary (oracle/_ref/pbrt, unmodified) rendering the same
.pbrt file on this box.  Sampler "halton" — the device evaluates the reference's own Halton points, so the two
films see identical samples.

Bars (floating point; written here): >= 95 % of the pixels within 1e-3 relative and relMSE <= 1e-4 — the bars of
the untextured halton tests; the measured figures are printed.  What differs: the last ulp of CUDA's logf
(MIPMap level) and sinf / cosf against glibc's, and the order in which a pixel's samples are summed.
"""
import numpy as np
import pytest

import texture_util as tu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(iile):
    if iile.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    iile.set_device(0)
    return iile


HEAD = '''LookAt 0.4 -2.6 1.6  0 0 0.1  0 0 1
Camera "perspective" "float fov" [42] %(camera)s
Sampler "halton" "integer pixelsamples" [%(spp)d]
Integrator "%(integrator)s" "integer maxdepth" [4]
Film "image" "integer xresolution" [160] "integer yresolution" [120] "string filename" "tex.pfm"
WorldBegin
LightSource "point" "point from" [1.5 -1.5 3] "color I" [14 14 14]
AttributeBegin
  AreaLightSource "diffuse" "color L" [6 6 6]
  Shape "trianglemesh" "point P" [-0.5 -0.5 2.5  0.5 -0.5 2.5  0.5 0.5 2.5  -0.5 0.5 2.5] "integer indices" [0 2 1 0 3 2]
AttributeEnd
'''

# a ground plane receding to the horizon (strong anisotropy: EWA with clamped eccentricity, many MIP levels)
# and a tilted quad with per-vertex normals (shading geometry: dndu / dndv for the bump map)
GEOM = '''%(textures)s
%(material)s
Shape "trianglemesh" "point P" [-6 -3 0  6 -3 0  6 30 0  -6 30 0] "integer indices" [0 1 2 0 2 3]
  "float uv" [0 0 4 0 4 11 0 11]
%(material2)s
Shape "trianglemesh" "point P" [-0.9 0.2 0.05  0.9 0.0 0.05  0.9 0.4 1.3  -0.9 0.6 1.3] "integer indices" [0 1 2 0 2 3]
  "float uv" [0 0 1 0 1 1 0 1] "normal N" [-0.3 -1 0.2  0.3 -1 0.2  0.3 -1 0.5  -0.3 -1 0.5]
WorldEnd
'''

CASES = {
    # name: (textures, material of the plane, material of the quad, integrator, camera params, spp)
    "kd_ewa": ('Texture "t" "spectrum" "imagemap" "string filename" "%(png)s"',
               'Material "matte" "texture Kd" "t"', 'Material "matte" "texture Kd" "t"', "path", "", 4),
    "kd_trilinear_clamp": ('Texture "t" "spectrum" "imagemap" "string filename" "%(png)s" "bool trilinear" ["true"] '
                           '"string wrap" "clamp" "float uscale" [0.5] "float vscale" [0.5] "float udelta" [0.1]',
                           'Material "matte" "texture Kd" "t"', 'Material "plastic" "texture Kd" "t" "float roughness" [0.2]',
                           "path", "", 4),
    "scale_tga_black": ('Texture "a" "spectrum" "imagemap" "string filename" "%(tga)s" "string wrap" "black" "float maxanisotropy" [2]\n'
                        'Texture "t" "spectrum" "scale" "texture tex1" "a" "color tex2" [0.9 0.5 0.4]',
                        'Material "matte" "texture Kd" "t"', 'Material "matte" "texture Kd" "a" "float sigma" [20]', "path", "", 4),
    "float_tex_sigma_rough": ('Texture "f" "float" "imagemap" "string filename" "%(pfm)s" "float scale" [40]\n'
                              'Texture "r" "float" "imagemap" "string filename" "%(png)s" "float scale" [0.4]',
                              'Material "matte" "color Kd" [.6 .6 .6] "texture sigma" "f"',
                              'Material "plastic" "color Kd" [.3 .3 .5] "color Ks" [.5 .5 .5] "texture roughness" "r"', "path", "", 4),
    "bump": ('Texture "b" "float" "imagemap" "string filename" "%(png)s" "float scale" [0.08]',
             'Material "matte" "color Kd" [.6 .6 .6] "texture bumpmap" "b"',
             'Material "plastic" "color Kd" [.5 .3 .3] "color Ks" [.4 .4 .4] "texture bumpmap" "b"', "path", "", 4),
    "bump_and_kd_dof": ('Texture "b" "float" "imagemap" "string filename" "%(png)s" "float scale" [0.05]\n'
                        'Texture "t" "spectrum" "imagemap" "string filename" "%(png)s"',
                        'Material "matte" "texture Kd" "t" "texture bumpmap" "b"', 'Material "uber" "texture Kd" "t" "color Ks" [.2 .2 .2]',
                        "path", '"float lensradius" [0.03] "float focaldistance" [3]', 4),
}


def agreement(got, want):
    d = np.abs(got - want)
    rel = (d / np.maximum(np.abs(want), 1e-2)).max(axis=-1)
    mse = ((got - want) ** 2).mean(axis=-1)
    relmse = float((mse / (want.mean(axis=-1) ** 2 + 1e-2)).mean())
    return float((rel < 1e-5).mean()), float((rel < 1e-3).mean()), relmse


@pytest.mark.parametrize("case", sorted(CASES))
def test_textured_film_matches_the_stock_binary(gpu, oracle, tmp_path, case):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    files = {"png": tmp_path / "a.png", "tga": tmp_path / "b.tga", "pfm": tmp_path / "c.pfm"}
    tu.write_png(files["png"], tu.test_image(37, 29, seed=3))
    tu.write_tga(files["tga"], tu.test_image(16, 24, seed=5))
    tu.write_pfm(files["pfm"], tu.test_image(20, 20, seed=7).astype(np.float32) / 255.0)
    textures, mat, mat2, integrator, camera, spp = CASES[case]
    text = HEAD % {"camera": camera, "spp": spp, "integrator": integrator} + GEOM % {
        "textures": textures % files, "material": mat, "material2": mat2}
    src = tmp_path / f"{case}.pbrt"
    src.write_text(text)
    scene = gpu.Scene.from_pbrt(src)
    film, stats = scene.render()
    rgb = scene.resolve(film)
    out = tmp_path / f"{case}_ref.pfm"
    oracle._run([oracle.REF_PBRT, "--quiet", "--outfile", out, src])
    ref = oracle.read_pfm(out)
    assert ref.shape == rgb.shape and ref.mean() > 0.01
    f5, f3, relmse = agreement(rgb, ref)
    print(f"{case}: {100 * f5:.2f} % of pixels within 1e-5, {100 * f3:.2f} % within 1e-3, relMSE {relmse:.3g}, "
          f"mean ratio {float(rgb.mean() / ref.mean()):.6f}")
    assert f3 >= 0.95 and relmse <= 1e-4, (f5, f3, relmse)
    assert abs(float(rgb.mean()) / float(ref.mean()) - 1) < 1e-3
    # the texture must matter: the same scene with the texture replaced by its mean is a different picture
    if case == "kd_ewa":
        flat = gpu.Scene.from_pbrt(_untextured(tmp_path, text))
        rgb0 = flat.resolve(flat.render()[0])
        assert np.abs(rgb0 - ref).mean() > 20 * np.abs(rgb - ref).mean()


def _untextured(tmp_path, text):
    f = tmp_path / "untextured.pbrt"
    f.write_text(text.replace('"texture Kd" "t"', '"color Kd" [.5 .5 .5]'))
    return f


def test_textured_scenes_are_refused_where_textures_are_not_evaluated(gpu, tmp_path):
    """Image textures and bump maps are evaluated by Integrator "path" on the main camera's rays; the direct
    integrators, the hemispheric pass and the IISPT task refuse such a scene instead of rendering it differently."""
    tu.write_png(tmp_path / "a.png", tu.test_image(8, 8))
    files = {"png": tmp_path / "a.png"}
    textures, mat, mat2, _, camera, spp = CASES["kd_ewa"]
    head = HEAD.replace('Sampler "halton"', 'Sampler "random"')
    text = head % {"camera": camera, "spp": spp, "integrator": "directlighting"} + GEOM % {
        "textures": textures % files, "material": mat, "material2": mat2}
    src = tmp_path / "direct.pbrt"
    src.write_text(text)
    scene = gpu.Scene.from_pbrt(src)
    with pytest.raises(gpu.IileError) as e:
        scene.render()
    assert "path integrator only" in str(e.value)
    src.write_text(text.replace('"directlighting"', '"path"'))
    scene = gpu.Scene.from_pbrt(src)
    with pytest.raises(gpu.IileError) as e:
        scene.render_hemispheres(np.float32([[0, 0, 1]]), np.float32([[0, 0, 1]]), size=8)
    assert "hemispheric pass" in str(e.value)
    with pytest.raises(gpu.IileError) as e:
        scene.iispt_hemi_points(gpu.IisptTask(0, 0, 16, 16, 8, 8, 0, 0, 0))
    assert "IISPT task" in str(e.value)
