#!/usr/bin/env python
"""Host-side load time of the cfg5 terrain (8M triangles): parse + parallel SAH build + upload.
python tools/time_build.py [threads ...]  (IILE_BUILD_TRACE prints the builder's phases on stderr)"""
import hashlib
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, time, hashlib
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_product
iile = load_product()
t = time.time()
s = iile.Scene.from_pbrt(sys.argv[1], upload=bool(iile.device_count()))
dt = time.time() - t
i = s.info()
s.save_flat("/tmp/_tb.flat")
print("threads=%%s load_s=%%.2f bvh_s=%%.2f nodes=%%d md5=%%s" %% (sys.argv[2], dt, i.buildSeconds, i.nNodes,
      hashlib.md5(open("/tmp/_tb.flat", "rb").read()).hexdigest()))
''' % (str(ROOT / "tests"), str(ROOT))


def main():
    sys.path.insert(0, str(ROOT))
    import bench
    path = bench.terrain_scene_path(0)
    for th in (sys.argv[1:] or ["1", str(os.cpu_count())]):
        env = dict(os.environ, IILE_BUILD_THREADS=th, IILE_BUILD_TRACE="1")
        r = subprocess.run([sys.executable, "-c", CHILD, str(path), th], env=env, capture_output=True, text=True)
        print(r.stdout.strip())
        print(r.stderr.strip())


if __name__ == "__main__":
    main()
