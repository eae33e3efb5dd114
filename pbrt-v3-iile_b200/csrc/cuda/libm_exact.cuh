// sinf / cosf with the bits of the host's libm.
//
// The reference calls std::sin / std::cos on floats, i.e. glibc's sinf / cosf (2.28 and later: the algorithm of Arm's
// optimized routines — argument reduction by pi/2 and an odd / even polynomial, all in double, rounded to float at
// the end).  Those are NOT correctly rounded: on 1.7 % / 0.9 % of the arguments ConcentricSampleDisk produces they
// return the other neighbour of the exact value (measured on this image's glibc 2.39, 30 M arguments), so neither
// CUDA's sinf (1-2 ulp) nor double-then-round (correctly rounded) reproduces the reference's bounce directions bit
// for bit.  This is a restatement of the published algorithm with its published polynomial coefficients; a C copy of
// it returns glibc's value on all 30,000,000 random arguments of [-pi/4, 2 pi] tested, with and without FMA
// contraction (the final rounding to float absorbs the difference), profiles/r9_ab_libm_double.txt; the fused
// form is used here (half the double-precision instructions).
// Valid for |x| < 120 (the callers pass angles of at most 2 pi); beyond that, and for NaN, double-then-round.
#pragma once

#include <cstdint>

namespace iile {

struct SincosPoly {
    double c0, c1, c2, c3, c4;  // cosine polynomial
    double s1, s2, s3;          // sine polynomial
};

__device__ __forceinline__ float sincosPolyEval(double x, double x2, bool negateCos, int n) {
    // (the second table entry of the published algorithm is the first with the cosine polynomial negated)
    const double sgn = negateCos ? -1.0 : 1.0;
    if ((n & 1) == 0) {
        const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
        const double x3 = x * x2;
        const double s1 = __fma_rn(x2, s3c, s2c);
        const double x7 = x3 * x2;
        const double s = __fma_rn(x3, s1c, x);
        return (float)__fma_rn(x7, s1, s);
    }
    const double c0 = sgn * 0x1p0, c1c = sgn * -0x1.ffffffd0c621cp-2, c2c = sgn * 0x1.55553e1068f19p-5,
                 c3c = sgn * -0x1.6c087e89a359dp-10, c4c = sgn * 0x1.99343027bf8c3p-16;
    const double x4 = x2 * x2;
    const double c2 = __fma_rn(x2, c4c, c3c);
    const double c1 = __fma_rn(x2, c2c, c1c);
    const double x6 = x4 * x2;
    const double c = __fma_rn(x2, c1, c0);
    return (float)__fma_rn(x6, c2, c);
}

__device__ __forceinline__ uint32_t absTop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ffu; }

// quadrant n and the reduced argument x - n * pi/2
__device__ __forceinline__ double reducePiOver2(double x, int *np) {
    const double r = x * 0x1.45F306DC9C883p+23;  // 2/pi * 2^24
    const int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
}

__device__ __forceinline__ float sinGlibc(float y) {
    double x = (double)y;
    if (absTop12(y) < absTop12(0x1.921FB6p-1f)) {  // |y| < pi/4
        if (absTop12(y) < absTop12(0x1p-12f)) return y;
        return sincosPolyEval(x, x * x, false, 0);
    }
    if (!(absTop12(y) < absTop12(120.0f))) return (float)sin(x);
    int n;
    x = reducePiOver2(x, &n);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign of the sine in quadrants 0..3: + - - +
    return sincosPolyEval(x * s, x * x, (n & 2) != 0, n);
}

__device__ __forceinline__ float cosGlibc(float y) {
    double x = (double)y;
    if (absTop12(y) < absTop12(0x1.921FB6p-1f)) {
        if (absTop12(y) < absTop12(0x1p-12f)) return 1.0f;
        return sincosPolyEval(x, x * x, false, 1);
    }
    if (!(absTop12(y) < absTop12(120.0f))) return (float)cos(x);
    int n;
    x = reducePiOver2(x, &n);
    const int m = n + 1;
    const double s = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
    return sincosPolyEval(x * s, x * x, (m & 2) != 0, n ^ 1);
}

// Both at once, without branches: one reduction, the sine polynomial and the cosine polynomial of the reduced
// argument, and the quadrant decides which is which and with what sign.  Bit for bit what sinGlibc / cosGlibc return:
// the published algorithm negates the ARGUMENT of the (odd) sine polynomial or every COEFFICIENT of the cosine
// polynomial, both of which negate the result exactly; for |y| < 2^-12 (where glibc returns y and 1) the
// polynomials round to the same values.
__device__ __forceinline__ void sincosGlibc(float y, float *sn, float *cs) {
    if (!(absTop12(y) < absTop12(120.0f))) {
        *sn = (float)sin((double)y);
        *cs = (float)cos((double)y);
        return;
    }
    int n;
    const double x = reducePiOver2((double)y, &n);
    const double x2 = x * x;
    const float sp = sincosPolyEval(x, x2, false, 0), cp = sincosPolyEval(x, x2, false, 1);
    const float a = (n & 1) ? cp : sp;   // |sin y|'s polynomial
    const float b = (n & 1) ? sp : cp;   // |cos y|'s polynomial
    *sn = (n & 2) ? -a : a;
    *cs = ((n + 1) & 2) ? -b : b;
}

}  // namespace iile
