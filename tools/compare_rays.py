#!/usr/bin/env python
"""Every ray of a render on the device (iile_render_dump_rays_dev) against every ray the reference traces for the same
scene and samples (ref_harness render with REF_LOG_RAYS: an aggregate that logs and forwards), as SETS of bit patterns
(origin and direction): how many of the device's rays does the reference not trace, by kind.  Camera rays are
identical; a bounce ray differs when anything on the way to its direction differs in the last bit — this tool found
the `wo = -ray.d` of the continuation and measured what glibc's sinf / cosf cost (DESIGN.md experiments).
python tools/compare_rays.py [scene.pbrt] [maxdepth] [spp]"""
import json
import os
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402


def compare(src0, depth=3, spp=4, explain=0):
    """-> {"device_rays", "reference_rays", kind: {"rays", "not_traced_by_the_reference"}}"""
    import torch
    src0 = Path(src0)
    iile = load_product()
    iile.set_device(0)
    work = Path("/tmp/compare_rays")
    work.mkdir(exist_ok=True)
    text = src0.read_text()
    text = re.sub(r'"integer pixelsamples" \[\d+\]', f'"integer pixelsamples" [{spp}]', text)
    text = re.sub(r'Integrator "path"[^\n]*', f'Integrator "path" "integer maxdepth" [{depth}]', text)
    src = src0.parent / "_compare_rays.pbrt" if "Include" in text else work / "s.pbrt"
    src.write_text(text)
    try:
        scene = iile.Scene.from_pbrt(src)
    finally:
        if src.parent != work:
            src.unlink()
    flat = work / "s.flat"
    scene.save_flat(flat)
    h, w = scene.film_shape()
    cap = 1 << 23
    rays_d = torch.empty(cap * 8, dtype=torch.float32, device="cuda")
    kinds_d = torch.empty(cap, dtype=torch.uint8, device="cuda")
    scratch = torch.zeros(h, w, 4, device="cuda")
    n = min(scene.render_dump_rays_dev(scratch.data_ptr(), iile.RenderParams(0, 0, 0, 0, 1, 0, 0), rays_d.data_ptr(),
                                       kinds_d.data_ptr(), cap, 1, 0), cap)
    mine = rays_d.view(cap, 8)[:n].cpu().numpy().copy()
    kinds = kinds_d[:n].cpu().numpy()
    os.environ["REF_LOG_RAYS"] = str(work / "ref_rays.bin")
    oracle.ref_render(flat, work / "ref.pfm", sampler="pixelstream", threads=1)
    del os.environ["REF_LOG_RAYS"]
    ref = np.fromfile(work / "ref_rays.bin", dtype="<f4").reshape(-1, 10)
    key = lambda a: {bytes(x) for x in np.ascontiguousarray(a[:, [0, 1, 2, 4, 5, 6]].astype("<f4"))}
    sr = key(ref)
    missing = np.array([bytes(x) not in sr for x in np.ascontiguousarray(mine[:, [0, 1, 2, 4, 5, 6]].astype("<f4"))])
    sm = key(mine)
    names = {0: "camera", 1: "bounce", 2: "mis", 3: "shadow"}
    out = {"scene": src0.name, "maxdepth": depth, "spp": spp, "device_rays": int(n), "reference_rays": int(len(ref)),
           "reference_rays_not_traced_by_the_device": int(sum(bytes(x) not in sm for x in
                                                             np.ascontiguousarray(ref[:, [0, 1, 2, 4, 5, 6]].astype("<f4"))))}
    for k in np.unique(kinds):
        out[names.get(int(k), str(k))] = {"rays": int((kinds == k).sum()), "not_traced_by_the_reference": int(missing[kinds == k].sum())}
    if explain:
        # each differing ray next to the reference ray closest to it (origin and direction, 6-D): which components
        # differ and by how many ulps — a last-bit difference in a direction points at a libm call, a gross one at a
        # branch taken differently
        from scipy.spatial import cKDTree
        cols = [0, 1, 2, 4, 5, 6]
        tree = cKDTree(ref[:, cols].astype("f8"))
        bad = np.nonzero(missing)[0]
        _, near = tree.query(mine[bad][:, cols].astype("f8"))
        ulps = lambda a, b: np.abs(a.astype("<f4").view("<i4").astype("i8") - b.astype("<f4").view("<i4").astype("i8"))
        for i, j in list(zip(bad, near))[:explain]:
            u = ulps(mine[i, cols], ref[j, cols])
            print("DIFF", names.get(int(kinds[i]), "?"), "o", mine[i, :3].tolist(), "d", mine[i, 4:7].tolist(),
                  "w", [float(mine[i, 3]), float(mine[i, 7])], "ulps(o,d)", u.tolist(), "ref_extra", ref[j, [3, 7, 8, 9]].tolist())
    return out


def main():
    src0 = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "scenes" / "smoke.pbrt"
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    print(json.dumps(compare(src0, depth, spp, int(os.environ.get("COMPARE_RAYS_EXPLAIN", "0")))))


if __name__ == "__main__":
    main()
