"""The device's libm (csrc/cuda/libm_exact.cuh) against the host's glibc, on the CPU.

The reference calls std::sin / std::cos / std::acos / std::atan2 / std::log on floats, i.e. glibc's sinf / cosf / acosf /
atan2f / logf, none of which is correctly rounded; the device restates glibc's algorithms so that a bounce direction, a
sphere's phi, an environment-map lookup or a MIP level carries the reference's bits (tests/test_rays_identical_gpu.py
shows the effect on the GPU).  Here the SAME header is compiled for the host (tests/libm_host: a shim of five CUDA
intrinsics) and compared with the host's libm, bit for bit, on a few million arguments of the ranges the callers
produce plus the awkward ones (zeros, subnormals, infinities, NaN, exponent extremes).  Bar: not one bit differs.
This pins the restatements' constants and control flow; nvcc's code generation of IEEE add / mul / div / sqrt / fma
is what the GPU test adds."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def libm(tmp_path_factory):
    out = tmp_path_factory.mktemp("libm") / "libm_exact_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fno-builtin",
                           "-I", str(ROOT / "pbrt-v3-iile_b200" / "csrc" / "cuda"),
                           str(ROOT / "tests" / "libm_host" / "libm_exact_host.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.libm_exact_mismatches.restype = C.c_long
    lib.libm_exact_mismatches.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long)]

    def mismatches(which, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float32)
        first = C.c_long(-1)
        bad = lib.libm_exact_mismatches(which, x.ctypes.data, None if y is None else y.ctypes.data, len(x), C.byref(first))
        where = None if first.value < 0 else (float(x[first.value]), None if y is None else float(y[first.value]))
        return bad, where
    return mismatches


SPECIAL = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-45, -1e-45, 1e-38, 1.1754944e-38, 3.4028235e38, -3.4028235e38,
                    np.inf, -np.inf, np.nan, 2.0 ** -12, 2.0 ** -29, 0.4375, 1.1875, 2.4375, 2.0 ** 26, 0.785398, 0.7853982,
                    119.99999, 120.0, 1e6, 6.2831855, 3.1415927, 1.5707964], dtype=np.float32)


@pytest.mark.parametrize("which,name", [(0, "sinf"), (1, "cosf"), (2, "sincosf")])
def test_sin_cos_have_glibc_bits(libm, which, name):
    rng = np.random.default_rng(11 + which)
    x = np.concatenate([rng.uniform(-0.8, 6.3, 3_000_000), rng.uniform(-120, 120, 1_000_000),
                        rng.uniform(-1, 1, 200_000) * 2.0 ** rng.integers(-40, 1, 200_000), SPECIAL]).astype(np.float32)
    assert libm(which, x) == (0, None), name


def test_acos_has_glibc_bits(libm):
    rng = np.random.default_rng(21)
    x = np.concatenate([rng.uniform(-1, 1, 3_000_000), 1 - 2.0 ** rng.uniform(-24, -1, 300_000),
                        -1 + 2.0 ** rng.uniform(-24, -1, 300_000), rng.uniform(-1, 1, 200_000) * 2.0 ** rng.integers(-60, 1, 200_000),
                        rng.uniform(-3, 3, 1000), SPECIAL]).astype(np.float32)
    assert libm(3, x) == (0, None)


def test_log_has_glibc_bits(libm):
    rng = np.random.default_rng(31)
    bits = rng.integers(0, 0x7F800000, 4_000_000, dtype=np.uint32)          # every positive finite float equally likely
    x = np.concatenate([bits.view(np.float32), rng.uniform(0, 1, 1_000_000).astype(np.float32),   # roughness values
                        rng.uniform(0.5, 2, 500_000).astype(np.float32), SPECIAL])
    assert libm(4, x) == (0, None)


def test_atan2_has_glibc_bits(libm):
    rng = np.random.default_rng(41)
    n = 3_000_000
    y, x = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    spread = slice(n // 2, n)                                                # exponents of both spread over 2^-40 .. 2^40
    y[spread] *= 2.0 ** rng.integers(-40, 41, n - n // 2)
    x[spread] *= 2.0 ** rng.integers(-40, 41, n - n // 2)
    gy, gx = np.meshgrid(SPECIAL, SPECIAL)
    # atan2f(y, 1) is atanf(y): the kernel on its own, every binade (glibc's large-argument bound is 2^25, not fdlibm's 2^26)
    ay = (rng.uniform(1, 2, 400_000) * 2.0 ** rng.integers(-40, 41, 400_000) * rng.choice([-1, 1], 400_000))
    y = np.concatenate([y, gy.ravel(), np.ones(1000), ay]).astype(np.float32)
    x = np.concatenate([x, gx.ravel(), rng.uniform(-4, 4, 1000), np.ones(len(ay))]).astype(np.float32)
    assert libm(5, y, x) == (0, None)
