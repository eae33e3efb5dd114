#!/usr/bin/env python
"""Render one scene a few times on cuda:0 and print device-timed Msamples/s (profiling driver for
ncu launch lists).  python tools/render_once.py cfg1|cfg3|cfg5|<file.pbrt> [spp] [reps]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    from conftest import load_product
    import oracle
    iile = load_product()
    name = sys.argv[1]
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    if name == "cfg5":
        import bench
        path = bench.terrain_scene_path(0)
    else:
        path = oracle.REF_SCENES / f"{name}.pbrt" if not name.endswith(".pbrt") else Path(name)
    iile.set_device(0)
    scene = iile.Scene.from_pbrt(path, upload=True)
    setup = scene.info().setup
    if spp:
        setup.pixelSamples = spp
        scene.set_setup(setup)
    for _ in range(reps):
        film, stats = scene.render()
    print(json.dumps({"scene": name, "spp": setup.pixelSamples, "device_s": stats.deviceSeconds,
                      "msamples_per_s": stats.cameraRays / stats.deviceSeconds / 1e6,
                      "rays": stats.closestRays + stats.shadowRays, "launches": stats.launches}))


if __name__ == "__main__":
    main()
