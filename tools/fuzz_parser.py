#!/usr/bin/env python
"""Mutated scene files through the front-end (iile_scene_load_pbrt, no upload): whatever the text, the call must
come back with a scene or an error code — no crash, no hang, no exception across the C boundary (INTEGRATION.md
"Error behaviour").  Mutations: tokens deleted / duplicated / swapped, brackets unbalanced, numbers replaced by huge,
negative, NaN or non-numeric ones, index lists corrupted, files truncated.  Each batch runs in a child process so that
a crash is caught and the offending file kept.  With `flat` as third argument the same for binary flat scene files
(iile_scene_load_flat): random bytes / 32-bit words overwritten (counts, indices, node offsets), truncation.
`ply` / `image`: mutated binary and ASCII PLY meshes (Shape "plymesh") and mutated PNG / TGA / PFM texture files
(Texture "imagemap") behind a fixed scene.  CPU only.  python tools/fuzz_parser.py [n] [seed] [flat|ply|image]"""
import json
import random
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WORK = Path("/tmp/fuzz_parser")

CHILD = r'''
import sys
from pathlib import Path
ROOT = Path(sys.argv[1]); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT))
from conftest import load_product
iile = load_product()
ok = err = 0
for f in sys.argv[2:]:
    print("FILE", f, flush=True)
    try:
        s = iile.Scene.from_flat(f, upload=False) if f.endswith(".flat") else iile.Scene.from_pbrt(f, upload=False)
        s.info(); s.close(); ok += 1
    except iile.IileError:
        err += 1
print("DONE", ok, err, flush=True)
'''


def mutate(rng, text):
    toks = text.split(" ")
    for _ in range(rng.choice((1, 1, 2, 4))):
        if not toks:
            break
        k = rng.randrange(len(toks))
        op = rng.randrange(9)
        if op == 0:
            del toks[k]
        elif op == 1:
            toks.insert(k, toks[k])
        elif op == 2:
            j = rng.randrange(len(toks))
            toks[k], toks[j] = toks[j], toks[k]
        elif op == 3:
            toks[k] = rng.choice(["[", "]", '"', "[[", "]]", '""', "#"])
        elif op == 4:
            toks[k] = rng.choice(["1e38", "-1e38", "nan", "inf", "-inf", "1e-45", "99999999999", "-1", "0", "2147483648", "x1"])
        elif op == 5:
            toks[k] = toks[k][: max(0, len(toks[k]) // 2)]
        elif op == 6:
            toks[k] = toks[k] + rng.choice(['"', "]", "[", "\n", "\x00", "\xff"])
        elif op == 7:
            toks = toks[: max(1, rng.randrange(len(toks)))]
        else:
            toks[k] = rng.choice(['"integer indices"', '"point P"', '"float fov"', '"string filename"', 'Shape', 'AttributeEnd',
                                  'WorldEnd', 'WorldBegin', 'Include', 'Texture', 'ObjectBegin', 'Material', 'AreaLightSource'])
    return " ".join(toks)


def mutate_flat(rng, data):
    b = bytearray(data)
    for _ in range(rng.choice((1, 2, 4, 16))):
        if len(b) < 16:
            break
        op = rng.randrange(4)
        if op == 0:
            b[rng.randrange(len(b))] = rng.randrange(256)
        elif op == 1:  # a 32-bit word: counts in the header, indices, offsets
            k = rng.randrange(0, max(1, len(b) - 4)) & ~3
            if rng.random() < 0.5:
                k = rng.randrange(0, min(len(b) - 4, 600)) & ~3
            b[k:k + 4] = rng.choice([b"\xff\xff\xff\xff", b"\x00\x00\x00\x80", b"\xff\xff\xff\x7f", b"\x00\x00\x00\x00",
                                     b"\x00\x00\x80\x7f", b"\x00\x00\xc0\x7f", bytes(rng.randrange(256) for _ in range(4))])
        elif op == 2:
            b = b[: rng.randrange(1, len(b))]
        else:
            k = rng.randrange(len(b))
            b[k:k] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 64)))
    return bytes(b)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    mode = sys.argv[3] if len(sys.argv) > 3 else "pbrt"
    flat = mode == "flat"
    WORK.mkdir(exist_ok=True)
    asset_seeds = []
    if mode in ("ply", "image"):
        sys.path.insert(0, str(ROOT / "tests"))
        import numpy as np
        import texture_util as tu
        if mode == "ply":
            v = np.random.default_rng(0).random((40, 3)).astype("<f4")
            idx = np.random.default_rng(1).integers(0, 40, (30, 3)).astype("<i4")
            head = ("ply\nformat binary_little_endian 1.0\nelement vertex 40\nproperty float x\nproperty float y\nproperty float z\n"
                    "element face 30\nproperty list uchar int vertex_indices\nend_header\n").encode()
            body = v.tobytes() + b"".join(b"\x03" + r.tobytes() for r in idx)
            asset_seeds.append(("m.ply", head + body))
            asc = ("ply\nformat ascii 1.0\nelement vertex 40\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\n"
                   "property float ny\nproperty float nz\nelement face 30\nproperty list uchar int vertex_indices\nend_header\n")
            asc += "".join("%g %g %g 0 0 1\n" % tuple(r) for r in v) + "".join("3 %d %d %d\n" % tuple(r) for r in idx)
            asset_seeds.append(("m.ply", asc.encode()))
        else:
            tu.write_png(WORK / "seed.png", tu.test_image(19, 13, seed=1))
            tu.write_tga(WORK / "seed.tga", tu.test_image(9, 7, seed=2))
            tu.write_pfm(WORK / "seed.pfm", tu.test_image(6, 5, seed=3).astype(np.float32) / 255)
            for nm in ("seed.png", "seed.tga", "seed.pfm"):
                asset_seeds.append(("t." + nm.split(".")[1], (WORK / nm).read_bytes()))
    seeds = [(ROOT / "scenes" / "smoke.pbrt").read_text()]
    tex = ROOT / "tests" / "test_textures.py"
    seeds.append('LookAt 0 -3 1 0 0 0 0 0 1\nCamera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [4]\n'
                 'Film "image" "integer xresolution" [32] "integer yresolution" [32]\nWorldBegin\n'
                 'Texture "t" "spectrum" "imagemap" "string filename" "missing.png" "float uscale" [2]\n'
                 'Texture "s" "spectrum" "scale" "texture tex1" "t" "color tex2" [.5 .5 .5]\n'
                 'LightSource "spot" "point from" [0 0 3] "point to" [0 0 0]\nMaterial "uber" "texture Kd" "s" "float index" [1.3]\n'
                 'Shape "plymesh" "string filename" "missing.ply"\nShape "loopsubdiv" "integer nlevels" [2] "integer indices" [0 1 2 0 2 3] '
                 '"point P" [0 0 0 1 0 0 1 1 0 0 1 0]\nShape "cylinder" "float radius" [.2]\nWorldEnd\n')
    flat_seeds = [(ROOT / "tests" / "golden" / "smoke.flat").read_bytes()] if flat else []
    crashes = done = total_ok = total_err = 0
    batch = 100
    while done < n:
        files = []
        for k in range(min(batch, n - done)):
            if asset_seeds:
                nm, data = rng.choice(asset_seeds)
                d = WORK / f"a_{k}"
                d.mkdir(exist_ok=True)
                (d / nm).write_bytes(mutate_flat(rng, data) if rng.random() < 0.9 else data)
                f = d / "s.pbrt"
                what = ('Shape "plymesh" "string filename" "m.ply"' if mode == "ply" else
                        'Texture "t" "spectrum" "imagemap" "string filename" "%s"\nMaterial "matte" "texture Kd" "t"\n'
                        'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]' % nm)
                f.write_text('LookAt 0 -3 1 0 0 0 0 0 1\nCamera "perspective"\nFilm "image" "integer xresolution" [8] "integer yresolution" [8]\n'
                             'WorldBegin\nLightSource "point"\nMaterial "matte"\n' + what + '\nWorldEnd\n')
            elif flat:
                f = WORK / f"m_{k}.flat"
                f.write_bytes(mutate_flat(rng, rng.choice(flat_seeds)))
            else:
                f = WORK / f"m_{k}.pbrt"
                f.write_bytes(mutate(rng, rng.choice(seeds)).encode("utf-8", "surrogateescape"))
            files.append(str(f))
        p = subprocess.run([sys.executable, "-c", CHILD, str(ROOT)] + files, capture_output=True, text=True, timeout=600)
        lines = p.stdout.splitlines()
        if lines and lines[-1].startswith("DONE"):
            _, a, b = lines[-1].split()
            total_ok += int(a)
            total_err += int(b)
        else:
            crashes += 1
            last = [l for l in lines if l.startswith("FILE")]
            culprit = last[-1].split(" ", 1)[1] if last else "?"
            keep = WORK / (f"crash_{crashes}" + (".flat" if flat else ".pbrt"))
            if culprit != "?":
                keep.write_bytes(Path(culprit).read_bytes())
                if asset_seeds:
                    import shutil
                    shutil.copytree(Path(culprit).parent, WORK / f"crash_{crashes}_dir", dirs_exist_ok=True)
            print("CRASH rc", p.returncode, "file kept as", keep, (p.stderr or "")[-300:])
        done += len(files)
    print(json.dumps({"files": done, "loaded": total_ok, "refused": total_err, "crashed_batches": crashes}))


if __name__ == "__main__":
    main()
