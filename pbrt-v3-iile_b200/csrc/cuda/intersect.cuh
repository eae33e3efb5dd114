// Device-side ray / box and ray / triangle tests.
//
// Arithmetic follows the reference operation for operation so that hit /
// miss decisions, t and barycentrics are bit-identical to the CPU renderer
// (which is built without FMA contraction):
//   slab test      Bounds3::IntersectP(ray, invDir, dirIsNeg)   src/core/geometry.h:1412-1438
//   triangle test  Triangle::Intersect / IntersectP             src/shapes/triangle.cpp:188-291, 405-505
// This translation unit MUST be compiled with -fmad=false: every a*b+c below
// is a rounded multiply followed by a rounded add.  Division and sqrt are
// IEEE-correct in CUDA by default (-prec-div / -prec-sqrt).
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace iile {

#define IILE_MACHINE_EPS 5.9604644775390625e-08f /* 2^-24, pbrt.h:198 */
__device__ __forceinline__ constexpr float gammaF(int n) {
    return (n * IILE_MACHINE_EPS) / (1 - n * IILE_MACHINE_EPS);
}

// Per-ray quantities that the reference recomputes per node / per triangle
// but that depend on the ray alone (hoisting them is exact).
#ifndef IILE_PICK_MASKS
#define IILE_PICK_MASKS 1
#endif
struct RayCtx {
    float ox, oy, oz;
    float dx, dy, dz;
    float tMax;
    float ix, iy, iz;   // 1/d
    int nx, ny, nz;     // dirIsNeg
    int negMask;        // nx | ny << 1 | nz << 2; bit 3 is always clear (chunk nodes rely on it, c_api.cu)
    int kx, ky, kz;     // triangle.cpp:205-209 permutation
#if IILE_PICK_MASKS
    unsigned m0, m1;    // all ones when kz == 0 / kz == 1: the permutation as two bit selects per component
#endif
    float Sx, Sy, Sz;   // triangle.cpp:216-218 shear
    float pdx, pdy, pdz;  // permuted direction (only z is used after the shear)
};

// IILE_PICK_MASKS: the permutation (kx, ky, kz) is a rotation fixed by kz, so the nine picks of a
// triangle test become bit selects on two per-ray masks (one LOP3 each, no compares, no predicates):
// profiles/r5y showed the compare-and-select form at 10 % of the closest-hit kernel's instructions.
__device__ __forceinline__ float bitsel(unsigned m, float a, float b) {
    return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m));
}

__device__ __forceinline__ float pick(float x, float y, float z, int k) {
    return k == 0 ? x : (k == 1 ? y : z);
}

__device__ __forceinline__ void initRay(RayCtx &r, float ox, float oy, float oz, float dx, float dy,
                                        float dz, float tMax) {
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.tMax = tMax;
    r.ix = 1.f / dx; r.iy = 1.f / dy; r.iz = 1.f / dz;
    r.nx = r.ix < 0; r.ny = r.iy < 0; r.nz = r.iz < 0;
    r.negMask = r.nx | (r.ny << 1) | (r.nz << 2);
    float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    // MaxDimension, geometry.h:1049-1053
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    float px = pick(dx, dy, dz, r.kx), py = pick(dx, dy, dz, r.ky), pz = pick(dx, dy, dz, r.kz);
    r.Sx = -px / pz;
    r.Sy = -py / pz;
    r.Sz = pick(r.ix, r.iy, r.iz, r.kz);  // 1.f / pz, already computed above
#if IILE_PICK_MASKS
    r.m0 = r.kz == 0 ? 0xffffffffu : 0u;
    r.m1 = r.kz == 1 ? 0xffffffffu : 0u;
#endif
}

// geometry.h:1412-1438.  n0 = {bmin.xyz, bmax.x}, n1 = {bmax.y, bmax.z, ...}.
// Returns the slab result WITHOUT the `tMin < ray.tMax` term, which the caller
// applies (it is the only place ray.tMax enters); *tMinOut is the entry t.
__device__ __forceinline__ bool slabTest(const float4 n0, const float4 n1, const RayCtx &r,
                                         float *tMinOut) {
    const float k = 1 + 2 * gammaF(3);
    float tMin = ((r.nx ? n0.w : n0.x) - r.ox) * r.ix;
    float tMax = ((r.nx ? n0.x : n0.w) - r.ox) * r.ix;
    float tyMin = ((r.ny ? n1.x : n0.y) - r.oy) * r.iy;
    float tyMax = ((r.ny ? n0.y : n1.x) - r.oy) * r.iy;
    tMax *= k;
    tyMax *= k;
    if (tMin > tyMax || tyMin > tMax) return false;
    if (tyMin > tMin) tMin = tyMin;
    if (tyMax < tMax) tMax = tyMax;
    float tzMin = ((r.nz ? n1.y : n0.z) - r.oz) * r.iz;
    float tzMax = ((r.nz ? n0.z : n1.y) - r.oz) * r.iz;
    tzMax *= k;
    if (tMin > tzMax || tzMin > tMax) return false;
    if (tzMin > tMin) tMin = tzMin;
    if (tzMax < tMax) tMax = tzMax;
    *tMinOut = tMin;
    return tMax > 0;
}

// triangle.cpp:188-291.  On a hit writes t and the three barycentrics.
__device__ __forceinline__ bool triangleTest(const RayCtx &r, float tMaxNow, float3 p0, float3 p1,
                                             float3 p2, float *tOut, float *b0Out, float *b1Out,
                                             float *b2Out) {
    // translate, permute
    float a0x = p0.x - r.ox, a0y = p0.y - r.oy, a0z = p0.z - r.oz;
    float a1x = p1.x - r.ox, a1y = p1.y - r.oy, a1z = p1.z - r.oz;
    float a2x = p2.x - r.ox, a2y = p2.y - r.oy, a2z = p2.z - r.oz;
#if IILE_PICK_MASKS
    // kz = 0: (kx, ky, kz) = (1, 2, 0) -> (y, z, x); kz = 1: (2, 0, 1) -> (z, x, y); kz = 2: (x, y, z)
    float p0x = bitsel(r.m0, a0y, bitsel(r.m1, a0z, a0x)), p0y = bitsel(r.m0, a0z, bitsel(r.m1, a0x, a0y)),
          p0z = bitsel(r.m0, a0x, bitsel(r.m1, a0y, a0z));
    float p1x = bitsel(r.m0, a1y, bitsel(r.m1, a1z, a1x)), p1y = bitsel(r.m0, a1z, bitsel(r.m1, a1x, a1y)),
          p1z = bitsel(r.m0, a1x, bitsel(r.m1, a1y, a1z));
    float p2x = bitsel(r.m0, a2y, bitsel(r.m1, a2z, a2x)), p2y = bitsel(r.m0, a2z, bitsel(r.m1, a2x, a2y)),
          p2z = bitsel(r.m0, a2x, bitsel(r.m1, a2y, a2z));
#else
    float p0x = pick(a0x, a0y, a0z, r.kx), p0y = pick(a0x, a0y, a0z, r.ky), p0z = pick(a0x, a0y, a0z, r.kz);
    float p1x = pick(a1x, a1y, a1z, r.kx), p1y = pick(a1x, a1y, a1z, r.ky), p1z = pick(a1x, a1y, a1z, r.kz);
    float p2x = pick(a2x, a2y, a2z, r.kx), p2y = pick(a2x, a2y, a2z, r.ky), p2z = pick(a2x, a2y, a2z, r.kz);
#endif
    // shear
    p0x += r.Sx * p0z; p0y += r.Sy * p0z;
    p1x += r.Sx * p1z; p1y += r.Sy * p1z;
    p2x += r.Sx * p2z; p2y += r.Sy * p2z;
    // edge functions
    float e0 = p1x * p2y - p1y * p2x;
    float e1 = p2x * p0y - p2y * p0x;
    float e2 = p0x * p1y - p0y * p1x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        double p2txp1ty = (double)p2x * (double)p1y;
        double p2typ1tx = (double)p2y * (double)p1x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0x * (double)p2y;
        double p0typ2tx = (double)p0y * (double)p2x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1x * (double)p0y;
        double p1typ0tx = (double)p1y * (double)p0x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
    float det = e0 + e1 + e2;
    if (det == 0) return false;
    p0z *= r.Sz;
    p1z *= r.Sz;
    p2z *= r.Sz;
    float tScaled = e0 * p0z + e1 * p1z + e2 * p2z;
    if (det < 0 && (tScaled >= 0 || tScaled < tMaxNow * det)) return false;
    else if (det > 0 && (tScaled <= 0 || tScaled > tMaxNow * det)) return false;
    float invDet = 1 / det;
    float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
    float t = tScaled * invDet;
    // conservative t > 0 check, triangle.cpp:272-291
    float maxZt = fmaxf(fmaxf(fabsf(p0z), fabsf(p1z)), fabsf(p2z));
    float deltaZ = gammaF(3) * maxZt;
    float maxXt = fmaxf(fmaxf(fabsf(p0x), fabsf(p1x)), fabsf(p2x));
    float maxYt = fmaxf(fmaxf(fabsf(p0y), fabsf(p1y)), fabsf(p2y));
    float deltaX = gammaF(5) * (maxXt + maxZt);
    float deltaY = gammaF(5) * (maxYt + maxZt);
    float deltaE = 2 * (gammaF(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
    float maxE = fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fabsf(e2));
    float deltaT = 3 * (gammaF(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
    if (t <= deltaT) return false;
    *tOut = t;
    *b0Out = b0;
    *b1Out = b1;
    *b2Out = b2;
    return true;
}

}  // namespace iile
