"""C-ABI surface and host logic (no GPU): the library loads, exports every symbol
include/iile_b200.h declares, fails loudly without a device, parses / rejects scenes like
the reference's front-end, and writes images the reference's readers accept."""
import re
import struct

import numpy as np
import pytest

from conftest import ROOT

GOLD = ROOT / "tests" / "golden"


def test_exports_every_declared_symbol(iile):
    header = (ROOT / "include" / "iile_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(iile_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(iile.lib, name), f"{name} declared in include/iile_b200.h but not exported"
    assert sorted(iile.EXPORTED_SYMBOLS) == declared
    assert iile.lib.iile_abi_version() == 4


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md's entry-point table maps each C-ABI symbol to the reference interface it replaces."""
    header = (ROOT / "include" / "iile_b200.h").read_text()
    doc = (ROOT / "INTEGRATION.md").read_text()
    # the table writes families compactly: `iile_comm_*`, `iile_dev_malloc/free/...`, `iile_intersect_closest(_dev)`
    expanded = doc
    for fam in re.findall(r"`(iile_[a-z_]+?)_?\*`", doc):
        expanded += " " + " ".join(n for n in re.findall(r"\b(iile_[a-z_0-9]+)\s*\(", header) if n.startswith(fam))
    for stem, rest in re.findall(r"`(iile_[a-z]+_)([a-z]+(?:/[a-z]+)+)`", doc):
        expanded += " " + " ".join(stem + r for r in rest.split("/"))
    for base in re.findall(r"`(iile_[a-z_]+)\(_dev\)`", doc):
        expanded += f" {base} {base}_dev"
    for name in sorted(set(re.findall(r"\b(iile_[a-z_0-9]+)\s*\(", header))):
        assert re.search(rf"\b{name}\b", expanded), f"{name} is not documented in INTEGRATION.md"


def test_struct_sizes_match_header(iile):
    import ctypes as C
    assert C.sizeof(iile.RenderSetup) == 520 and iile.RenderSetup.cameraType.offset == 500
    assert iile.RAY_DTYPE.itemsize == 32 and iile.HIT_DTYPE.itemsize == 16
    assert iile.NODE_DTYPE.itemsize == 32 and iile.PRIM_DTYPE.itemsize == 32


def test_no_cpu_fallback(iile):
    """Without a CUDA device every compute entry point must fail, never compute on the host."""
    if iile.device_count() > 0:
        pytest.skip("a GPU is present")
    s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
    rays = np.zeros(4, dtype=iile.RAY_DTYPE)
    rays["d"][:, 2] = 1
    with pytest.raises(iile.IileError) as e:
        s.intersect(rays)
    assert e.value.code == -3
    with pytest.raises(iile.IileError):
        s.render()
    with pytest.raises(iile.IileError):
        iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=True)


def test_smoke_scene_flattening(iile, tmp_path):
    s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
    i = s.info()
    assert (i.nPrims, i.nSpheres, i.nLights, i.nMaterials) == (29, 1, 3, 8)
    assert (i.setup.xRes, i.setup.yRes, i.setup.pixelSamples, i.setup.maxDepth) == (96, 64, 8, 6)
    assert list(i.sampleBounds) == [0, 0, 96, 64] and list(i.croppedBounds) == [0, 0, 96, 64]
    # 3 lights, more than one => default "spatial" strategy (lightdistrib.cpp:47-65)
    assert i.setup.lightSampleStrategy == 2
    p = tmp_path / "s.flat"
    s.save_flat(p)
    t = iile.Scene.from_flat(p, upload=False)
    assert t.nodes().tobytes() == s.nodes().tobytes()
    assert np.array_equal(t.prim_order(), s.prim_order())


@pytest.mark.parametrize("body,needle", [
    ('Shape "cone"', "not supported"),
    ('ObjectBegin "a"', "instancing"),
    ('Material "metal"', "not supported"),
    ('Material "mix"', "not supported"),
    ('LightSource "infinite" "string mapname" "sky.exr"', "not supported"),
    ('LightSource "goniometric"', "not supported"),
    ('MakeNamedMedium "fog"', "media"),
    ('Shape "trianglemesh" "point P" [0 0 0 1 0 0 0 1 0] "integer indices" [0 1 7]', "out-of-bounds"),
    ('Frobnicate', "unknown directive"),
    ('WorldEnd\nIntegrator "bdpt"\nWorldBegin', "not supported"),
    ('WorldEnd\nCamera "realistic"\nWorldBegin', "not supported"),
])
def test_unsupported_directives_are_errors(iile, tmp_path, body, needle):
    """anything outside the supported subset must raise, not silently differ."""
    f = tmp_path / "bad.pbrt"
    f.write_text('Sampler "random"\nWorldBegin\nShape "sphere"\n' + body + '\nShape "sphere"\nWorldEnd\n')
    with pytest.raises(iile.IileError) as e:
        iile.Scene.from_pbrt(f, upload=False)
    assert e.value.code == -2 and needle in str(e.value)


def test_default_sampler_is_halton(iile, tmp_path):
    # the reference defaults to halton, 16 spp (api.cpp:168-179, halton.cpp:124-130)
    f = tmp_path / "h.pbrt"
    f.write_text('WorldBegin\nShape "sphere"\nWorldEnd\n')
    s = iile.Scene.from_pbrt(f, upload=False)
    setup = s.info().setup
    assert setup.samplerType == iile.SAMPLER_HALTON and setup.pixelSamples == 16 and setup.samplePixelCenter == 0
    f.write_text('Sampler "halton" "integer pixelsamples" [8] "bool samplepixelcenter" ["true"]\n'
                 'WorldBegin\nShape "sphere"\nWorldEnd\n')
    setup = iile.Scene.from_pbrt(f, upload=False).info().setup
    assert setup.samplerType == iile.SAMPLER_HALTON and setup.pixelSamples == 8 and setup.samplePixelCenter == 1
    f.write_text('Sampler "random"\nWorldBegin\nShape "sphere"\nWorldEnd\n')
    setup = iile.Scene.from_pbrt(f, upload=False).info().setup
    assert setup.samplerType == iile.SAMPLER_RANDOM and setup.pixelSamples == 4
    # sobol.cpp:64-70 + sobol.h:64-70: default 16, rounded up to a power of two
    f.write_text('Sampler "sobol" "integer pixelsamples" [6]\nWorldBegin\nShape "sphere"\nWorldEnd\n')
    setup = iile.Scene.from_pbrt(f, upload=False).info().setup
    assert setup.samplerType == iile.SAMPLER_SOBOL and setup.pixelSamples == 8
    f.write_text('Sampler "stratified"\nWorldBegin\nShape "sphere"\nWorldEnd\n')
    with pytest.raises(iile.IileError) as e:
        iile.Scene.from_pbrt(f, upload=False)
    assert e.value.code == -2 and "stratified" in str(e.value)


def test_path_samples_override_installs_sobol(iile, tmp_path, monkeypatch):
    """IILE: CreatePathIntegrator replaces the sampler by a SobolSampler of IILE_PATH_SAMPLES_OVERRIDE
    samples per pixel (path.cpp:201-212)."""
    f = tmp_path / "o.pbrt"
    f.write_text('Sampler "random" "integer pixelsamples" [4]\nIntegrator "path"\nWorldBegin\nShape "sphere"\nWorldEnd\n')
    monkeypatch.setenv("IILE_PATH_SAMPLES_OVERRIDE", "24")
    setup = iile.Scene.from_pbrt(f, upload=False).info().setup
    assert setup.samplerType == iile.SAMPLER_SOBOL and setup.pixelSamples == 32
    monkeypatch.delenv("IILE_PATH_SAMPLES_OVERRIDE")
    setup = iile.Scene.from_pbrt(f, upload=False).info().setup
    assert setup.samplerType == iile.SAMPLER_RANDOM and setup.pixelSamples == 4


def test_old_flat_files_read_as_random_sampler(iile):
    # samplerType / samplePixelCenter took the last 8 bytes of filename[256]: the golden flat
    # scene predates them and must still load, as the random sampler
    from conftest import ROOT
    s = iile.Scene.from_flat(ROOT / "tests" / "golden" / "smoke.flat", upload=False)
    assert s.info().setup.samplerType == iile.SAMPLER_RANDOM


def test_crop_window_and_filter_bounds(iile, tmp_path):
    f = tmp_path / "c.pbrt"
    f.write_text('Film "image" "integer xresolution" [200] "integer yresolution" [100] '
                 '"float cropwindow" [.25 .75 .1 .9]\nPixelFilter "gaussian"\nSampler "random"\n'
                 'WorldBegin\nShape "sphere"\nWorldEnd\n')
    i = iile.Scene.from_pbrt(f, upload=False).info()
    assert list(i.croppedBounds) == [50, 10, 150, 90]          # film.cpp:55-59
    assert list(i.sampleBounds) == [48, 8, 152, 92]            # film.cpp:77-83, radius 2
    assert i.setup.filterType == 1 and i.setup.filterRadius[0] == 2.0


def test_pfm_and_exr_writers(iile, tmp_path):
    s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
    rng = np.random.default_rng(1)
    rgb = rng.random((64, 96, 3), dtype=np.float32) * 4
    s.write_image(tmp_path / "a.pfm", rgb)
    raw = (tmp_path / "a.pfm").read_bytes()
    assert raw.startswith(b"PF\n96 64\n-1.000000\n")            # imageio.cpp:450-458
    back = iile.read_pfm(tmp_path / "a.pfm")
    assert np.array_equal(back, rgb)
    s.write_image(tmp_path / "a.exr", rgb)
    exr = (tmp_path / "a.exr").read_bytes()
    assert exr[:4] == bytes([0x76, 0x2f, 0x31, 0x01])
    # parse our own header enough to find the pixel data and decode the half R channel
    pos = 8
    attrs = {}
    while exr[pos] != 0:
        e = exr.index(b"\0", pos); name = exr[pos:e].decode(); pos = e + 1
        e = exr.index(b"\0", pos); typ = exr[pos:e].decode(); pos = e + 1
        size = struct.unpack_from("<i", exr, pos)[0]; pos += 4
        attrs[name] = (typ, exr[pos:pos + size]); pos += size
    pos += 1
    assert struct.unpack("<4i", attrs["dataWindow"][1]) == (0, 0, 95, 63)
    assert struct.unpack("<4i", attrs["displayWindow"][1]) == (0, 0, 95, 63)
    offsets = struct.unpack_from("<64Q", exr, pos)
    y, nbytes = struct.unpack_from("<2i", exr, offsets[5])
    assert (y, nbytes) == (5, 96 * 8)
    line = np.frombuffer(exr, dtype="<f2", count=96 * 4, offset=offsets[5] + 8).reshape(4, 96)
    assert np.array_equal(line[3], rgb[5, :, 0].astype(np.float16))   # R, round-to-nearest-even
    assert np.array_equal(line[2], rgb[5, :, 1].astype(np.float16))   # G
    assert np.array_equal(line[1], rgb[5, :, 2].astype(np.float16))   # B
    assert np.all(line[0] == 1)                                       # A


def test_film_resolve_arithmetic(iile):
    """to_rgb_array (film.cpp:187-222): RGB->XYZ->RGB, divide by weight, clamp at 0, scale."""
    s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
    film = np.zeros((64, 96, 4), np.float32)
    film[..., :3] = np.float32([2.0, 4.0, 8.0])
    film[..., 3] = 4.0
    film[0, 0] = [-1, -1, -1, 2]      # negative sums clamp to 0
    film[0, 1] = [1, 1, 1, 0]         # zero weight: left unnormalised
    rgb = s.resolve(film)
    assert np.allclose(rgb[5, 5], [0.5, 1.0, 2.0], rtol=1e-5)
    assert np.all(rgb[0, 0] == 0)
    assert np.allclose(rgb[0, 1], 1.0, rtol=1e-5)


def test_scene_create_from_caller_arrays(iile, oracle):
    """iile_scene_create: the flattened scene as caller-owned arrays — what a host that
    keeps the reference's parser hands over.  Without nodes the BVH is built here (same tree as the
    flat file's); with the caller's own LinearBVHNode array + primitive order nothing is rebuilt."""
    flat = oracle.FlatScene(GOLD / "smoke.flat")            # an IILEFLT1 file: 64-byte materials, 40-byte lights
    ref = iile.Scene.from_flat(GOLD / "smoke.flat", upload=False)
    setup = ref.info().setup

    def widen(raw, width):
        out = np.zeros((len(raw), width), np.uint8)
        out[:, :raw.shape[1]] = raw
        return out

    mats, lights = widen(flat.materials, iile.MATERIAL_BYTES), widen(flat.lights, iile.LIGHT_BYTES)
    spheres = flat.spheres.view(np.uint8).reshape(len(flat.spheres), -1)
    common = dict(N=flat.N, UV=flat.UV, spheres=spheres, materials=mats, lights=lights, upload=False)
    built = iile.Scene.from_desc(flat.P, flat.prims, setup, **common)
    assert built.nodes().tobytes() == ref.nodes().tobytes()
    assert np.array_equal(built.prim_order(), ref.prim_order())
    assert built.info().buildSeconds > 0
    given = iile.Scene.from_desc(flat.P, flat.prims, setup, nodes=ref.nodes(), prim_order=ref.prim_order(), **common)
    assert given.nodes().tobytes() == ref.nodes().tobytes() and given.info().buildSeconds == 0
    i, j = given.info(), ref.info()
    assert (i.nTriangles, i.nSpheres, i.nMaterials, i.nLights) == (j.nTriangles, j.nSpheres, j.nMaterials, j.nLights)
    # malformed descriptions are errors, not crashes
    bad = flat.prims.copy()
    bad["v"][0, 0] = len(flat.P) + 5
    with pytest.raises(iile.IileError) as e:
        iile.Scene.from_desc(flat.P, bad, setup, **common)
    assert e.value.code == -1 and "out of range" in str(e.value)
    order = ref.prim_order().copy()
    order[3] = len(order) + 7
    with pytest.raises(iile.IileError):
        iile.Scene.from_desc(flat.P, flat.prims, setup, nodes=ref.nodes(), prim_order=order, **common)


def test_front_end_accepts_the_documented_subset(iile, tmp_path, oracle):
    """One scene using every directive family INTEGRATION.md lists as accepted; the flattened scene
    carries what the directives said (types, counts, constant-folded textures, light records)."""
    f = tmp_path / "all.pbrt"
    f.write_text('''
LookAt 0 -5 2  0 0 0  0 0 1
Camera "orthographic" "float screenwindow" [-2 2 -1.5 1.5] "float lensradius" [.01] "float focaldistance" [5]
Film "image" "integer xresolution" [40] "integer yresolution" [30] "float cropwindow" [.1 .9 .2 .8] "string filename" "all.exr"
PixelFilter "mitchell" "float B" [.3] "float C" [.35]
Sampler "random" "integer pixelsamples" [3]
Integrator "directlighting" "string strategy" "one" "integer maxdepth" [4]
Accelerator "bvh" "string splitmethod" "hlbvh" "integer maxnodeprims" [2]
WorldBegin
Texture "half" "float" "mix" "float tex1" [.2] "float tex2" [.6]
Texture "tint" "spectrum" "scale" "color tex1" [.5 .5 .5] "color tex2" [.4 .8 1]
MakeNamedMaterial "sub" "string type" "substrate" "texture Kd" "tint" "float uroughness" [.2]
LightSource "point" "color I" [2 2 2] "point from" [0 0 3]
LightSource "spot" "point from" [1 1 3] "point to" [0 0 0] "float coneangle" [40]
LightSource "distant" "point from" [0 0 1] "point to" [0 0 0] "color L" [.5 .5 .5]
AttributeBegin
  Rotate 30 1 0 0
  LightSource "infinite" "color L" [.1 .2 .3]
AttributeEnd
AttributeBegin
  AreaLightSource "diffuse" "color L" [4 4 4] "bool twosided" "true"
  Translate 0 0 2.5
  Shape "disk" "float radius" [.5]
AttributeEnd
Material "matte" "float sigma" [10]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-3 -3 0 3 -3 0 3 3 0 -3 3 0]
NamedMaterial "sub"
Shape "sphere" "float radius" [.4] "float zmax" [.2]
Material "translucent" "texture roughness" "half"
Shape "cylinder" "float radius" [.2] "float zmin" [0] "float zmax" [1]
Material "uber" "color Kr" [.1 .1 .1]
Shape "loopsubdiv" "integer nlevels" [1] "integer indices" [0 1 2 0 2 3 0 3 1 1 3 2]
      "point P" [1 1 1  2 1 1  1.5 2 1  1.5 1.5 2]
Material "glass" "float uroughness" [.1] "float vroughness" [.3]
Shape "trianglemesh" "integer indices" [0 1 2] "point P" [-1 -1 .5 -2 -1 .5 -1.5 -2 .5]
WorldEnd
''')
    s = iile.Scene.from_pbrt(f, upload=False)
    i = s.info()
    st = i.setup
    assert (st.cameraType, st.integratorType, st.directStrategy, st.maxDepth, st.splitMethod, st.maxNodePrims) == (1, 1, 1, 4, 1, 2)
    assert st.filterType == iile.FILTER_MITCHELL and st.pixelSamples == 3 and st.filename == b"all.exr"
    assert (i.nSpheres, i.nLights) == (3, 5)          # disk + sphere + cylinder share the quadric record
    assert tuple(i.croppedBounds) == (4, 6, 36, 24)
    s.save_flat(tmp_path / "all.flat")
    flat = oracle.FlatScene(tmp_path / "all.flat")
    mats = flat.materials.view("<i4")[:, 0]
    assert sorted(set(mats.tolist())) == [0, 3, 4, 5, 6]          # matte, glass, uber, substrate, translucent
    sub = flat.materials[mats == 5][0].view("<f4")
    assert np.allclose(sub[2:5], [.2, .4, .5])                    # Kd = tint = scale of two constants
    tr = flat.materials[mats == 6][0].view("<f4")
    assert np.isclose(tr[12], .4)                                 # roughness = mix(.2, .6, .5)
    kinds = flat.lights.view("<i4")[:, 0]
    assert kinds.tolist() == [0, 2, 3, 4, 1]                      # point, spot, distant, infinite, area (scene order)
    quad_flags = flat.spheres["flags"] if "flags" in (flat.spheres.dtype.names or ()) else None
    if quad_flags is not None:
        assert sorted((quad_flags & 12).tolist()) == [0, 4, 8]
    infinite = flat.lights[3].view("<f4")
    assert infinite[49] > 1 and infinite[25] == 1                 # worldRadius (float 49) filled from the scene bound; lightToWorld[15] = 1


def _corrupt_nodes(oracle, src, dst, edit):
    raw = bytearray(src.read_bytes())
    flat = oracle.FlatScene(src)
    n_nodes, n_prims = len(flat.nodes), len(flat.prims)
    off = len(raw) - 4 * n_prims - 32 * n_nodes
    nodes = np.frombuffer(raw, dtype=oracle.NODE_DTYPE, count=n_nodes, offset=off)   # a view: edits land in raw
    edit(nodes, n_prims)
    dst.write_bytes(bytes(raw))


@pytest.mark.parametrize("name", ["second child backwards", "second child out of range", "axis", "leaf range",
                                  "leaf overlap", "nan box"])
def test_flat_scene_with_a_broken_node_array_is_rejected(iile, oracle, tmp_path, name):
    """a caller-supplied LinearBVHNode array is validated before anything walks it."""
    src = GOLD / "smoke.flat"
    flat = oracle.FlatScene(src)
    interior = int(np.flatnonzero(flat.nodes["nPrims"] == 0)[1])
    leaves = np.flatnonzero(flat.nodes["nPrims"] > 0)

    def edit(nodes, n_prims):
        if name == "second child backwards":
            nodes["offset"][interior] = interior          # a cycle for a walker that trusts it
        elif name == "second child out of range":
            nodes["offset"][interior] = len(nodes) + 5
        elif name == "axis":
            nodes["axis"][interior] = 7
        elif name == "leaf range":
            nodes["offset"][leaves[-1]] = n_prims - 1
            nodes["nPrims"][leaves[-1]] = 3
        elif name == "leaf overlap":
            nodes["offset"][leaves[1]] = nodes["offset"][leaves[0]]
        elif name == "nan box":
            nodes["bmin"][interior, 1] = np.nan

    bad = tmp_path / "bad.flat"
    _corrupt_nodes(oracle, src, bad, edit)
    with pytest.raises(iile.IileError) as e:
        iile.Scene.from_flat(bad, upload=False)
    assert "BVH" in str(e.value) or "primitive slot" in str(e.value)
    good = tmp_path / "good.flat"
    _corrupt_nodes(oracle, src, good, lambda nodes, n: None)
    iile.Scene.from_flat(good, upload=False).close()


def test_exr_is_read_back_by_openexr(iile, tmp_path):
    """the product's EXR bytes pinned by an independent reader — OpenEXR 2.3 as bundled in
    OpenCV (cv2.imread), not the test's own header parser.  What the reference's WriteImageEXR promises
    (imageio.cpp:164-189; ImageIO.RoundTripEXR, src/tests/imageio.cpp:83-89): half RGBA, pixel (x, y) of
    the data window = the float rounded to binary16, data window = the cropped pixel bounds inside the
    full-resolution display window."""
    import os
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    s = iile.Scene.from_pbrt(ROOT / "scenes" / "smoke.pbrt", upload=False)
    rng = np.random.default_rng(7)
    # values across the half range: tiny, ordinary, large (the film scale of a bright light), zero
    rgb = (rng.random((37, 53, 3), dtype=np.float32) * np.float32(8)) ** np.float32(4)
    rgb[0, 0] = 0
    rgb[1, 1] = [6.1e-5, 1e-7, 65504.0]          # smallest normal half, a subnormal, the largest half
    for name, (full, off) in {"full.exr": ((53, 37), (0, 0)), "crop.exr": ((200, 100), (50, 10))}.items():
        path = tmp_path / name
        iile._check(iile.lib.iile_write_image(os.fsencode(str(path)), rgb.ctypes.data_as(iile.C.c_void_p), 53, 37,
                                              full[0], full[1], off[0], off[1]))
        img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
        assert img is not None, "OpenEXR could not read the file"
        assert img.dtype == np.float32 and img.shape[:2] == (37, 53)           # the data window
        back = img[..., [2, 1, 0]]                                              # OpenCV hands out B, G, R(, A)
        want = rgb.astype(np.float16).astype(np.float32)
        assert np.array_equal(back, want), name
        if img.shape[2] == 4:
            assert np.all(img[..., 3] == 1.0)                                   # Rgba(r, g, b): a = 1
        # the windows, from the header (OpenCV only exposes the data window's size)
        raw = path.read_bytes()
        at = raw.index(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0") + 4
        assert struct.unpack_from("<4i", raw, at) == (off[0], off[1], off[0] + 52, off[1] + 36)
        at = raw.index(b"displayWindow\0box2i\0") + len(b"displayWindow\0box2i\0") + 4
        assert struct.unpack_from("<4i", raw, at) == (0, 0, full[0] - 1, full[1] - 1)
        # and a file OpenEXR itself writes from the same pixels decodes to the same values
        ref_path = tmp_path / ("ocv_" + name)
        assert cv2.imwrite(str(ref_path), rgb[..., ::-1].copy(), [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF])
        ref = cv2.imread(str(ref_path), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(ref[..., :3], img[..., :3])
