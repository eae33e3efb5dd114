#!/usr/bin/env python
"""Mutated scene files that the front-end ACCEPTS, rendered: whatever survives parsing (NaN colours, degenerate
cameras, zero-area lights, inverted crop windows that are still legal, huge intensities, ...) must render to a film
without a CUDA error, a hang or a crash — finite or not is the scene's business, the reference scrubs NaN / inf
samples itself (integrator.cpp:293-314) and so does the film kernel.  Children run batches under a time limit; a
batch that dies or times out is reported with its file.  Needs a GPU.  python tools/fuzz_mutant_renders.py [n] [seed]"""
import json
import random
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
from fuzz_parser import mutate  # noqa: E402

WORK = Path("/tmp/fuzz_mutants")

CHILD = r'''
import sys
from pathlib import Path
import numpy as np
ROOT = Path(sys.argv[1]); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT))
from conftest import load_product
iile = load_product(); iile.set_device(0)
rendered = refused = skipped = nonfinite = 0
for f in sys.argv[2:]:
    print("FILE", f, flush=True)
    try:
        s = iile.Scene.from_pbrt(f)
        i = s.info().setup
        h, w = s.film_shape()
        if h * w > (1 << 16) or h * w * i.pixelSamples > (1 << 21) or i.maxDepth > 64:
            skipped += 1; s.close(); continue
        film, st = s.render()
        rgb = s.resolve(film)
        nonfinite += int(not np.isfinite(rgb).all())
        rendered += 1
        s.close()
    except iile.IileError as e:
        if "CUDA" in str(e) or "cuda" in str(e):
            print("CUDAERR", f, str(e)[:200], flush=True)
            sys.exit(3)
        refused += 1
print("DONE", rendered, refused, skipped, nonfinite, flush=True)
'''


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    WORK.mkdir(exist_ok=True)
    seed_text = (ROOT / "scenes" / "smoke.pbrt").read_text().replace('"integer xresolution" [96] "integer yresolution" [64]',
                                                                        '"integer xresolution" [48] "integer yresolution" [32]')
    tot = [0, 0, 0, 0]
    bad = done = 0
    while done < n:
        files = []
        for k in range(min(50, n - done)):
            f = WORK / f"m_{k}.pbrt"
            f.write_bytes(mutate(rng, seed_text).encode("utf-8", "surrogateescape"))
            files.append(str(f))
        try:
            p = subprocess.run([sys.executable, "-c", CHILD, str(ROOT)] + files, capture_output=True, text=True, timeout=240)
            lines = p.stdout.splitlines()
            okay = bool(lines) and lines[-1].startswith("DONE")
        except subprocess.TimeoutExpired as e:
            lines = (e.stdout or b"").decode(errors="replace").splitlines() if isinstance(e.stdout, bytes) else (e.stdout or "").splitlines()
            okay = False
            p = None
        if okay:
            for j, v in enumerate(lines[-1].split()[1:]):
                tot[j] += int(v)
        else:
            bad += 1
            last = [l for l in lines if l.startswith("FILE")]
            culprit = last[-1].split(" ", 1)[1] if last else "?"
            if culprit != "?":
                (WORK / f"bad_{bad}.pbrt").write_bytes(Path(culprit).read_bytes())
            print("BAD BATCH", "timeout" if p is None else f"rc {p.returncode}", culprit, [l for l in lines if l.startswith("CUDAERR")][:1],
                  ((p.stderr if p else "") or "")[-300:])
        done += len(files)
    print(json.dumps({"files": done, "rendered": tot[0], "refused": tot[1], "skipped_large": tot[2], "films_with_non_finite_pixels": tot[3],
                      "bad_batches": bad}))


if __name__ == "__main__":
    main()
