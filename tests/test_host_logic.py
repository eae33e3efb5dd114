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
