#!/usr/bin/env python
"""The device's libm header (csrc/cuda/libm_exact.cuh, compiled for the host as tests/test_libm_exact.py does) against
the host's glibc on EVERY float of the ranges its callers produce: sinf / cosf / sincosf on (-120, 120) (the restatement's domain), acosf on
[-1, 1], logf on every positive finite float, atanf (atan2f's kernel) on every float.  CPU only, a few minutes.  python tools/libm_copies/exhaustive.py"""
import ctypes as C
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent


def main():
    out = Path(tempfile.mkdtemp()) / "libm_exact_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fno-builtin",
                           "-I", str(ROOT / "pbrt-v3-iile_b200" / "csrc" / "cuda"),
                           str(ROOT / "tests" / "libm_host" / "libm_exact_host.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.libm_exact_mismatches.restype = C.c_long
    lib.libm_exact_mismatches.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
    two_pi = int(np.float32(120.0).view(np.uint32)) - 1  # the restatement's whole domain (callers stay within 2 pi)
    one = 0x3F800000
    jobs = [("sinf", 0, two_pi, True), ("cosf", 1, two_pi, True), ("sincosf", 2, two_pi, True), ("acosf", 3, one, True),
            ("logf", 4, 0x7F800000 - 1, False), ("atanf = atan2f(y, 1)", 5, 0x7F800000, True)]
    chunk = 1 << 24

    def run(which, lo, hi, negate):
        bits = np.arange(lo, hi, dtype=np.uint32)
        if negate:
            bits |= np.uint32(0x80000000)
        first = C.c_long(-1)
        ones = np.ones(len(bits), dtype=np.float32) if which == 5 else None
        return lib.libm_exact_mismatches(which, bits.ctypes.data, None if ones is None else ones.ctypes.data, len(bits),
                                         C.byref(first))  # ctypes drops the GIL

    for name, which, top, both_signs in jobs:
        spans = [(lo, min(lo + chunk, top + 1), neg) for neg in ([False, True] if both_signs else [False])
                 for lo in range(0, top + 1, chunk)]
        with ThreadPoolExecutor(8) as ex:
            bad = sum(ex.map(lambda s: run(which, *s), spans))
        n = sum(hi - lo for lo, hi, _ in spans)
        print(f"{name}: {bad} mismatches on {n} floats (every float of {'+-' if both_signs else ''}[0, {np.uint32(top).view(np.float32)}])", flush=True)


if __name__ == "__main__":
    sys.exit(main())
