#!/usr/bin/env python
"""Builds pbrt-v3-iile_b200/data/sobol_matrices.bin from the reference's constant tables.

The Sobol' generator matrices are published data (Joe & Kuo direction numbers "new-joe-kuo-6.21201",
packed by L. Gruenschloss; the reference carries them as C arrays in src/core/sobolmatrices.cpp).  They
cannot be regenerated without that data set, so the numbers — not the source file — are extracted once,
where /root/reference exists, into a binary blob that is committed and linked into libiile_b200.so
(csrc/host/sobol_tables.cpp).  Layout, little endian:

  u32 magic 'SOBL', u32 nDims (1024), u32 matrixSize (52), u32 nVdC (25), u32 nVdCInv (26), u32 pad[3]
  u32 SobolMatrices32[nDims * matrixSize]           (SobolSampleFloat, lowdiscrepancy.h:259-272)
  u64 VdCSobolMatrices[nVdC][matrixSize]            (SobolIntervalToIndex, lowdiscrepancy.h:229-249)
  u64 VdCSobolMatricesInv[nVdCInv][matrixSize]

  python tools/extract_sobol_tables.py [/root/reference/src/core/sobolmatrices.cpp]
"""
import re
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def table(text, name):
    m = re.search(r"\b" + name + r"\s*\[[^=]*=\s*\{(.*?)\};", text, re.S)
    body = re.sub(r"/\*.*?\*/|//[^\n]*", "", m.group(1), flags=re.S)
    return [int(tok.rstrip("uUlL"), 0) for tok in re.findall(r"0[xX][0-9a-fA-F]+[uUlL]*|\d+[uUlL]*", body)]


def rows(text, name, size):
    """A 2-D table whose rows are partially initialised ({a, b}, the rest zero): padded to `size`."""
    m = re.search(r"\b" + name + r"\s*\[\]\s*\[[^=]*=\s*\{(.*?)\}\s*;", text, re.S)
    body = re.sub(r"/\*.*?\*/|//[^\n]*", "", m.group(1), flags=re.S)
    out = []
    for row in re.findall(r"\{([^{}]*)\}", body):
        vals = [int(tok.rstrip("uUlL"), 0) for tok in re.findall(r"0[xX][0-9a-fA-F]+[uUlL]*|\d+[uUlL]*", row)]
        assert len(vals) <= size
        out.append(vals + [0] * (size - len(vals)))
    return out


def main():
    src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/core/sobolmatrices.cpp")
    text = src.read_text()
    m32 = np.array(table(text, "SobolMatrices32"), dtype="<u4")
    n_dims, size = 1024, 52
    vdc = np.array(rows(text, "VdCSobolMatrices", size), dtype="<u8").reshape(-1)
    inv = np.array(rows(text, "VdCSobolMatricesInv", size), dtype="<u8").reshape(-1)
    assert len(m32) == n_dims * size, len(m32)
    assert len(vdc) % size == 0 and len(inv) % size == 0 and len(inv) >= len(vdc), (len(vdc), len(inv))
    out = ROOT / "pbrt-v3-iile_b200" / "data" / "sobol_matrices.bin"
    with open(out, "wb") as f:
        f.write(struct.pack("<4sIIIIIII", b"SOBL", n_dims, size, len(vdc) // size, len(inv) // size, 0, 0, 0))
        f.write(m32.tobytes())
        f.write(vdc.tobytes())
        f.write(inv.tobytes())
    print(out, out.stat().st_size, "bytes;", len(vdc) // size, "+", len(inv) // size, "VdC rows")


if __name__ == "__main__":
    main()
