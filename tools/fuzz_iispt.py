#!/usr/bin/env python
"""The IILE-specific passes over the material / light / shape variants of tests/test_variants_gpu.py: for random
variants of scenes/smoke.pbrt, (1) a batch of hemispheres at first-hit points against the reference's RenderView
(ref_harness hemi) with tests/test_hemi_gpu.py's bars, (2) one whole-film IISPT task against the reference's
find_intersection / compute_fpixel_weights / sample_hemisphere handed the device's films (ref_harness iispt_task) with
tests/test_iispt_task_gpu.py's bars for scenes with specular surfaces.  With `textures` as first argument the same
two checks over EVERY textured scene of tests/test_textures_gpu.py, the reference side reading the .pbrt file itself
(its own parser builds the textures).  python tools/fuzz_iispt.py [n] [seed] | textures"""
import json
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402
import test_variants_gpu as tv  # noqa: E402
from test_hemi_gpu import check as hemi_check  # noqa: E402

# variants that change what the IILE passes see (materials, lights, shapes); integrator / film / sampler edits do not
KEEP = [k for k in sorted(tv.CASES) if tv.CASES[k][0] in (tv.SPHERE, tv.AREA, tv.PLASTIC, tv.MATTE_SIGMA, tv.GLASS, tv.MIRROR,
                                                         tv.POINT, tv.RED_WALL, tv.ACCEL) and len(tv.CASES[k]) == 2]


def textured():
    import test_textures_gpu as tt
    iile = load_product()
    iile.set_device(0)
    work = Path("/tmp/fuzz_iispt_tex")
    work.mkdir(exist_ok=True)
    fails = 0
    pos0 = np.float32([[0.3, -1.0, 0.05], [2.0, 9.0, 0.02], [0.0, -0.6, 0.7], [-1.5, 3.0, 0.4], [0.2, 0.0, 1.9]])
    dir0 = np.float32([[0, 0, 1], [0, 0, 1], [0, 1, 0.1], [0.5, 0.5, -0.2], [0, 0.2, -1]])
    dir0 /= np.linalg.norm(dir0, axis=1, keepdims=True)
    for case in sorted(tt.CASES):
        src = tt._write_case(work, case, sampler="random")
        scene = iile.Scene.from_pbrt(src)
        problems = []
        try:
            want_i, want_n, want_z, _ = oracle.ref_hemispheres(src, pos0, dir0, work, spp=2)
            inten, nrm, dist, _ = scene.render_hemispheres(pos0, dir0, size=32, spp=2)
            frac, relmse = hemi_check(inten, nrm, dist, want_i, want_n, want_z)
            hemi_txt = f"hemi {frac:.4f} {relmse:.2g}"
        except AssertionError as e:
            problems.append("hemispheres: " + str(e)[:80])
            hemi_txt = "hemi FAIL"
        x0, y0, x1, y1, ts = 30, 20, 130, 110, 12
        task = iile.IisptTask(x0, y0, x1, y1, ts, 32, 77, 1, 99)
        samples, pos, direction, valid, films = scene.iispt_task(task, lambda x: x[:, 0:3])
        hemi, _, pixels, _ = oracle.ref_iispt_task(src, work, x0, y0, x1, y1, ts, 32, 77, 1, 99, films=films.cpu().numpy())
        keep = valid.astype(bool)
        if not np.array_equal(hemi[:, 0] > 0, keep):
            problems.append("hemi validity")
        same = float((samples[..., 3] == pixels[..., 3]).mean())
        took = (pixels[..., 3] > 0) & (samples[..., 3] > 0)
        got, want = samples[..., :3][took], pixels[..., :3][took]
        rel = (np.abs(got - want) / np.maximum(np.abs(want), 1e-3)).max(axis=-1)
        f4, f2 = float((rel < 1e-4).mean()), float((rel < 1e-2).mean())
        ratio = float(got.mean() / want.mean()) if want.size and want.mean() > 0 else 1.0
        if same < 0.995 or f2 < 0.9 or abs(ratio - 1) > 2e-2:
            problems.append(f"gather same={same:.4f} f2={f2:.4f} ratio={ratio:.4f}")
        fails += bool(problems)
        print("FAIL" if problems else "ok  ", case, f"{int(took.sum())} samples, {100 * f4:.2f} % within 1e-4, same-pixels {same:.4f}, "
              f"ratio {ratio:.5f}; {hemi_txt}", problems or "")
    print(json.dumps({"textured_scenes": len(tt.CASES), "failed": fails}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "textures":
        return textured()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    iile = load_product()
    iile.set_device(0)
    work = Path("/tmp/fuzz_iispt")
    work.mkdir(exist_ok=True)
    fails = 0
    for k in range(n):
        picks = rng.sample(KEEP, rng.choice((1, 2)))
        text, ok = tv.SMOKE, True
        for name in picks:
            old, new = tv.CASES[name][:2]
            if old not in text:
                ok = False
                break
            text = text.replace(old, new, 1)
        if not ok:
            continue
        src = work / "s.pbrt"
        src.write_text(text)
        scene = iile.Scene.from_pbrt(src)
        flat = work / "s.flat"
        scene.save_flat(flat)
        problems = []
        # ---- (2) one task over the whole 96 x 64 film ------------------------------------------------------------------
        x0, y0, x1, y1, ts = 0, 0, 96, 64, rng.choice((6, 9, 16))
        seed = rng.randrange(1 << 30)
        task = iile.IisptTask(x0, y0, x1, y1, ts, 32, 500, 2, seed)
        samples, pos, direction, valid, films = scene.iispt_task(task, lambda x: x[:, 0:3])
        hemi, _, pixels, info = oracle.ref_iispt_task(flat, work, x0, y0, x1, y1, ts, 32, 500, 2, seed, films=films.cpu().numpy())
        keep = valid.astype(bool)
        if not np.array_equal(hemi[:, 0] > 0, keep):
            problems.append("hemi validity")
        else:
            close = np.isclose(hemi[keep, 1:4], pos[keep], rtol=2e-5, atol=2e-3).all(axis=1).mean() if keep.any() else 1.0
            if close < 0.9:
                problems.append(f"hemi positions {close:.3f}")
        same = float((samples[..., 3] == pixels[..., 3]).mean())
        took = (pixels[..., 3] > 0) & (samples[..., 3] > 0)
        got, want = samples[..., :3][took], pixels[..., :3][took]
        rel = (np.abs(got - want) / np.maximum(np.abs(want), 1e-3)).max(axis=-1)
        f4, f2 = float((rel < 1e-4).mean()), float((rel < 1e-2).mean())
        ratio = float(got.mean() / want.mean()) if want.size and want.mean() > 0 else 1.0
        if same < 0.995 or f2 < 0.9 or abs(ratio - 1) > 2e-2:
            problems.append(f"gather same={same:.4f} f2={f2:.4f} ratio={ratio:.4f}")
        # ---- (1) hemispheres at the task's own hemi points ----------------------------------------------------------------
        hp, hd = pos[keep][:12], direction[keep][:12]
        try:
            want_i, want_n, want_z, _ = oracle.ref_hemispheres(flat, hp, hd, work)
            inten, nrm, dist, _ = scene.render_hemispheres(hp, hd)
            frac, relmse = hemi_check(inten, nrm, dist, want_i, want_n, want_z)
            hemi_txt = f"hemi {frac:.4f} {relmse:.2g}"
        except AssertionError as e:
            problems.append("hemispheres: " + str(e)[:80])
            hemi_txt = "hemi FAIL"
        if problems:
            fails += 1
            (work / f"fail_{fails}.pbrt").write_text(text)
        print("FAIL" if problems else "ok  ", picks, f"tile {ts}: {int(took.sum())} samples, {100 * f4:.2f} % within 1e-4, "
              f"same-pixels {same:.4f}, ratio {ratio:.5f}; {hemi_txt}", problems or "")
    print(json.dumps({"scenes": n, "failed": fails}))


if __name__ == "__main__":
    main()
