"""Image textures, host side (VERDICT r1 missing #1): the PNG / TGA / PFM readers, ImageTexture's
conversion and the MIPMap pyramid of csrc/host/texture_host.cpp against the UNMODIFIED reference's
ImageTexture::GetTexture (src/textures/imagemap.cpp:53-105, src/core/mipmap.h:100-237) run by
`ref_harness mipmap`.  Bar: every texel of every level bit-identical (same float operations in the same order
on the same host libm)."""
import numpy as np
import pytest

import texture_util as tu


SCENE = '''LookAt 0 0 3  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Sampler "halton" "integer pixelsamples" [4]
Film "image" "integer xresolution" [64] "integer yresolution" [64] "string filename" "t.pfm"
WorldBegin
LightSource "point" "point from" [0 0 3] "color I" [10 10 10]
Texture "img" "%s" "imagemap" "string filename" "%s" %s
Material "matte" "texture %s" "img"
Shape "trianglemesh" "point P" [-1 -1 0  1 -1 0  1 1 0  -1 1 0] "integer indices" [0 1 2 0 2 3]
  "float uv" [0 0 1 0 1 1 0 1]
WorldEnd
'''


@pytest.mark.parametrize("fmt,size,kind,wrap,extra", [
    ("png", (16, 16), "spectrum", "repeat", ""),
    ("png", (13, 7), "spectrum", "repeat", ""),                  # not powers of two: Lanczos resampling
    ("png", (21, 12), "spectrum", "clamp", '"float scale" [0.7]'),
    ("png", (9, 20), "float", "black", ""),                      # Texture<Float>: luminance, then gamma
    ("tga", (10, 6), "spectrum", "repeat", ""),
    ("pfm", (12, 10), "spectrum", "repeat", '"bool gamma" ["true"]'),
    ("pfm", (8, 8), "float", "clamp", ""),
])
def test_pyramid_matches_reference(iile, oracle, tmp_path, fmt, size, kind, wrap, extra):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    w, h = size
    img = tu.test_image(w, h, seed=w * 31 + h)
    path = tmp_path / f"tex.{fmt}"
    if fmt == "png":
        tu.write_png(path, img)
    elif fmt == "tga":
        tu.write_tga(path, img)
    else:
        tu.write_pfm(path, img.astype(np.float32) / 200.0)
    channels = 1 if kind == "float" else 3
    scene_file = tmp_path / "tex.pbrt"
    params = f'"string wrap" "{wrap}" {extra}'
    scene_file.write_text(SCENE % (kind, path, params, "sigma" if kind == "float" else "Kd"))
    s = iile.Scene.from_pbrt(scene_file, upload=False)
    flat = tmp_path / "tex.flat"
    s.save_flat(flat)
    tex, images, texels, mt = tu.read_flat_textures(flat)
    assert len(images) == 1 and int(images[0]["channels"]) == channels
    assert tex[-1]["type"] == 1 and mt.shape == (1, 12) and (mt >= 0).sum() == 1
    mine = tu.pyramid_levels(images[0], texels)
    scale = 0.7 if "scale" in extra else 1.0
    gamma = 1 if (fmt in ("png", "tga") or "gamma" in extra) else 0
    oracle._run([oracle.REF_HARNESS, "mipmap", path, channels, wrap, scale, gamma, tmp_path / "ref_pyr.bin"])
    ref = tu.read_ref_pyramid(tmp_path / "ref_pyr.bin", channels)
    assert len(ref) == len(mine)
    for l, (a, b) in enumerate(zip(mine, ref)):
        assert a.shape == b.shape, (l, a.shape, b.shape)
        assert np.array_equal(a.view("u4"), b.view("u4")), f"level {l} differs: max |d| = {np.abs(a - b).max()}"
    # a flat file with textures reloads to the same arrays
    s2 = iile.Scene.from_flat(flat, upload=False)
    s2.save_flat(tmp_path / "tex2.flat")
    assert (tmp_path / "tex2.flat").read_bytes() == flat.read_bytes()


def test_image_files_without_pixels_are_refused(iile, tmp_path):
    """Found by tools/fuzz_parser.py: a TGA header announcing a width of zero reached the MIPMap builder, which
    divided by it."""
    import texture_util as tu
    tu.write_tga(tmp_path / "a.tga", tu.test_image(8, 8))
    d = bytearray((tmp_path / "a.tga").read_bytes())
    d[12:14] = b"\x00\x00"
    (tmp_path / "a.tga").write_bytes(d)
    f = tmp_path / "s.pbrt"
    f.write_text(SCENE % ("spectrum", tmp_path / "a.tga", "", "Kd"))
    with pytest.raises(iile.IileError) as e:
        iile.Scene.from_pbrt(f, upload=False)
    assert "without pixels" in str(e.value)


def test_texture_errors(iile, tmp_path):
    """What the front-end does not implement is refused with a message, never rendered differently."""
    tu.write_png(tmp_path / "a.png", tu.test_image(8, 8))
    base = SCENE % ("spectrum", tmp_path / "a.png", "", "Kd")
    for bad, needle in ((base.replace('"string filename"', '"string mapping" "spherical" "string filename"'), "mapping"),
                        (base.replace("a.png", "a.exr"), ".png, .tga and .pfm"),
                        (base.replace('"string filename"', '"float maxanisotropy" [500] "string filename"'), "maxanisotropy"),
                        (base.replace("Material ", "".join('Texture "n%d" "spectrum" "scale" "texture tex1" "%s" "color tex2" [.9 .9 .9]\n'
                                                           % (k, "img" if k == 0 else "n%d" % (k - 1)) for k in range(4)) + "Material ", 1), "nested")):
        f = tmp_path / "bad.pbrt"
        f.write_text(bad)
        with pytest.raises(iile.IileError) as e:
            iile.Scene.from_pbrt(f, upload=False)
        assert needle in str(e.value), str(e.value)
