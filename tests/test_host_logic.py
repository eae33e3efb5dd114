"""Host-side logic of the Python mirror that runs without a GPU."""
def test_schedule_monitor_mirror():
    """IisptScheduleMonitor::next_task (iisptschedulemonitor.cpp:41-76), restated here line by line as the check: tasks of
    10 x 10 cells sweep the film row by row, the cell size shrinks by sqrt(0.79541357) (float) after every sweep."""
    import importlib
    import math
    import numpy as np
    iile = importlib.import_module("pbrt-v3-iile_b200")
    for bounds, start in (((0, 0, 96, 64), 6.0), ((-2, -2, 402, 402), 100.0), ((5, 7, 33, 300), 2.5)):
        m = iile.IisptScheduleMonitor(bounds, start)
        x0, y0, x1, y1 = bounds
        radius, nx, ny, pas = np.float32(start), x0, y0, 1
        mult = np.sqrt(np.float32(0.79541357))
        for k in range(60):
            eff = max(int(math.floor(radius)), 1)
            size = eff * 10
            want = {"x0": nx, "y0": ny, "x1": min(nx + size, x1), "y1": min(ny + size, y1), "tilesize": eff, "pass": pas,
                    "taskNumber": k}
            nx += size
            if nx >= x1:
                nx, ny = x0, ny + size
            if ny >= y1:
                ny, radius, pas = y0, np.float32(radius * mult), pas + 1
            assert m.next_task() == want


def test_non_finite_geometry_and_bad_setups_are_refused(tmp_path):
    """Found by tools/fuzz_parser.py: an infinite vertex coordinate crashed the SAH builder (bucket index of a NaN
    offset).  Geometry must be finite; a render setup is range-checked wherever it comes from."""
    import importlib
    import struct
    import pytest
    from conftest import ROOT
    iile = importlib.import_module("pbrt-v3-iile_b200")
    text = (ROOT / "scenes" / "smoke.pbrt").read_text()
    for bad in ("inf", "-inf", "nan"):
        f = tmp_path / "bad.pbrt"
        f.write_text(text.replace('"point P" [-.6 -.6 0  .6 -.6 0', f'"point P" [-.6 {bad} 0  .6 -.6 0', 1))
        with pytest.raises(iile.IileError) as e:
            iile.Scene.from_pbrt(f, upload=False)
        assert "not finite" in str(e.value) or "number" in str(e.value), str(e.value)
    f = tmp_path / "radius.pbrt"
    f.write_text(text.replace('Shape "sphere" "float radius" [.45]', 'Shape "sphere" "float radius" [inf]'))
    with pytest.raises(iile.IileError):
        iile.Scene.from_pbrt(f, upload=False)
    # a flat file whose setup is out of range: xRes sits 204 bytes into the file (8 magic + 32 counts + 164 of setup)
    data = bytearray((ROOT / "tests" / "golden" / "smoke.flat").read_bytes())
    assert struct.unpack_from("<ii", data, 204) == (96, 64)
    for value, needle in ((struct.pack("<i", -5), "resolution"), (struct.pack("<i", 1 << 30), "resolution")):
        d = bytearray(data)
        d[204:208] = value
        g = tmp_path / "bad.flat"
        g.write_bytes(d)
        with pytest.raises(iile.IileError) as e:
            iile.Scene.from_flat(g, upload=False)
        assert needle in str(e.value)
    scene = iile.Scene.from_flat(ROOT / "tests" / "golden" / "smoke.flat", upload=False)
    for field, value, needle in (("filterType", 9, "filter"), ("pixelSamples", 0, "pixelsamples"), ("integratorType", 7, "integrator"),
                                 ("samplerType", -1, "sampler"), ("cameraType", 3, "camera")):
        s = iile.RenderSetup.from_buffer_copy(scene.info().setup)
        setattr(s, field, value)
        with pytest.raises(iile.IileError) as e:
            scene.set_setup(s)
        assert needle in str(e.value), str(e.value)
    s = iile.RenderSetup.from_buffer_copy(scene.info().setup)
    s.filterRadius[0] = 0.0
    with pytest.raises(iile.IileError):
        scene.set_setup(s)
    s.filterRadius[0] = float("nan")
    with pytest.raises(iile.IileError):
        scene.set_setup(s)
    scene.set_setup(iile.RenderSetup.from_buffer_copy(scene.info().setup))  # an unchanged setup is accepted
