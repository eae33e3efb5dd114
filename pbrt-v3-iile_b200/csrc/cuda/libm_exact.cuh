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
// polynomials round to the same values (but for the sign of a zero, restored below).
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
    *sn = (y == 0.f) ? y : ((n & 2) ? -a : a);  // sinf(-0) is -0; the polynomial's sum would make it +0
    *cs = ((n + 1) & 2) ? -b : b;
}

// ---- acosf, atanf, atan2f: glibc 2.39 still carries the fdlibm float versions (e_acosf.c, s_atanf.c, e_atan2f.c:
// rational / polynomial approximations evaluated in float).  Restated here operation by operation — every step is an
// IEEE float add / multiply / divide / sqrt, which this library's -fmad=false build reproduces exactly; this header
// compiled for the host agrees with glibc on EVERY float for acosf ([-1, 1]) and atanf (tools/libm_copies/exhaustive.py); C copies
// return glibc's value on 30,000,000 of 30,000,000 random arguments (acosf) and 200,000,000 of 200,000,000 (atan2f, exponents
// of both arguments spread over 2^-40 ... 2^40).
__device__ __forceinline__ float acosGlibc(float x) {
    const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
                pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f;
    const float qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    const int32_t hx = __float_as_int(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    if (ix > 0x3f800000) return (x - x) / (x - x);
    float z, s = 0.f;
    const bool small = ix < 0x3f000000;  // |x| < 0.5
    if (small) {
        if (ix <= 0x23000000) return pio2_hi + pio2_lo;
        z = x * x;
    } else {
        z = (hx < 0 ? (one + x) : (one - x)) * 0.5f;
        s = sqrtf(z);
    }
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = p / q;
    if (small) return pio2_hi - (x - (pio2_lo - x * r));
    if (hx < 0) {
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    const float df = __int_as_float(__float_as_int(s) & 0xfffff000);
    const float c = (z - df * df) / (s + df);
    const float w = r * s + c;
    return 2.0f * (df + w);
}

__device__ __forceinline__ float atanGlibc(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const float one = 1.0f;
    const int32_t hx = __float_as_int(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25 (glibc's bound, found on the host; fdlibm's float version has 2^26 here)
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {  // |x| < 0.4375
        if (ix < 0x31000000) return x;  // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {  // |x| < 1.1875
            if (ix < 0x3f300000) {
                id = 0;
                x = (2.0f * x - one) / (2.0f + x);
            } else {
                id = 1;
                x = (x - one) / (x + one);
            }
        } else if (ix < 0x401c0000) {  // |x| < 2.4375
            id = 2;
            x = (x - 1.5f) / (one + 1.5f * x);
        } else {
            id = 3;
            x = -1.0f / x;
        }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    const float hi = id == 0 ? atanhi[0] : (id == 1 ? atanhi[1] : (id == 2 ? atanhi[2] : atanhi[3]));
    const float lo = id == 0 ? atanlo[0] : (id == 1 ? atanlo[1] : (id == 2 ? atanlo[2] : atanlo[3]));
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return hx < 0 ? -r : r;
}

__device__ __forceinline__ float atan2Glibc(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = __float_as_int(x), ix = hx & 0x7fffffff, hy = __float_as_int(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanGlibc(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);  // 2 * sign(x) + sign(y)
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000 || iy == 0x7f800000) return (float)atan2((double)y, (double)x);  // infinities: not reached by the callers
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 24) z = pi_o_2 + 0.5f * pi_lo;  // glibc's float thresholds (|y / x| beyond 2^24, below 2^-26), not fdlibm's 60:
    else if (hx < 0 && k < -26) z = 0.0f;   // found on the host — with 60 the copy missed glibc on 2 of 10^4 such arguments
    else z = atanGlibc(fabsf(y / x));
    if (m == 0) return z;
    if (m == 1) return __int_as_float(__float_as_int(z) ^ (int)0x80000000);
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

// ---- logf: glibc 2.27 and later (Arm's optimized routines): x = 2^k z, z in [0.7, 1.4) cut into 16 intervals with
// tabulated 1/c and log c, a cubic in r = z / c - 1, all in double, one rounding.  Correctly rounded on all but
// 416,909 of the 2,139,095,039 positive floats — on those double-then-round returns the other neighbour.  Table and
// coefficients are the published ones (they are also what this image's libm.so.6 holds); a C copy returns glibc's
// value on EVERY positive finite float, with and without FMA contraction.  Callers: RoughnessToAlpha, Log2 of the MIP
// level (pbrt.h:325-328).
static __device__ const double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

__device__ __forceinline__ float logGlibc(float x) {
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {  // zero, subnormal, negative, inf, NaN
        if (ix * 2 == 0) return -__int_as_float(0x7f800000);
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __int_as_float(0x7fc00000);
        ix = __float_as_uint(x * 0x1p23f);
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int32_t)tmp >> 23;
    const double z = (double)__uint_as_float(ix - (tmp & 0xff800000u));
    const double r = __fma_rn(z, kLogfTab[i][0], -1.0);
    const double y0 = __fma_rn((double)k, 0x1.62e42fefa39efp-1, kLogfTab[i][1]);
    const double r2 = r * r;
    double y = __fma_rn(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = __fma_rn(-0x1.00ea348b88334p-2, r2, y);
    return (float)__fma_rn(y, r2, y0 + r);
}

}  // namespace iile
