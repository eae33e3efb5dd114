// TEST INFRASTRUCTURE (tests/test_libm_exact.py): the device's libm restatements — the very header the kernels include,
// pbrt-v3-iile_b200/csrc/cuda/libm_exact.cuh — compiled for the HOST through a shim of the few CUDA intrinsics it uses,
// and compared with the host's glibc (what the reference calls through std::sin / std::cos / std::acos / std::atan2 /
// std::log on floats).  Built with -ffp-contract=off -fno-builtin so that neither side is folded or fused.
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline uint32_t __float_as_uint(float f) { uint32_t i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(uint32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }

#include "libm_exact.cuh"

static inline bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

// which: 0 sinf, 1 cosf, 2 sincos (both), 3 acosf, 4 logf, 5 atan2f(x[i], y[i]).  Returns the number of arguments on
// which the restatement's bits differ from glibc's; *firstBad = index of the first such argument (or -1).
extern "C" long libm_exact_mismatches(int which, const float *x, const float *y, long n, long *firstBad) {
    long bad = 0;
    *firstBad = -1;
    for (long i = 0; i < n; ++i) {
        volatile float a = x[i];
        bool ok = true;
        switch (which) {
        case 0: ok = same(iile::sinGlibc(a), sinf(a)); break;
        case 1: ok = same(iile::cosGlibc(a), cosf(a)); break;
        case 2: {
            float s, c;
            iile::sincosGlibc(a, &s, &c);
            ok = same(s, sinf(a)) && same(c, cosf(a));
            break;
        }
        case 3: ok = same(iile::acosGlibc(a), acosf(a)); break;
        case 4: ok = same(iile::logGlibc(a), logf(a)); break;
        case 5: {
            volatile float b = y[i];
            ok = same(iile::atan2Glibc(a, b), atan2f(a, b));
            break;
        }
        }
        if (!ok) {
            if (*firstBad < 0) *firstBad = i;
            ++bad;
        }
    }
    return bad;
}
