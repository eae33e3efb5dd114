// Device-side Shape "sphere": the quadric test with running error bounds.
//
// Restates Sphere::Intersect / IntersectP (src/shapes/sphere.cpp:49-130,
// 171-215) including the interval arithmetic of EFloat
// (src/core/efloat.h:48-283: every operation widens [low, high] by one ulp
// each way) and the ray transform with error terms
// (src/core/transform.h:287-312, 337-395).  The sphere stays in object space,
// as in the reference; only the hit decision and t are produced here — the
// shading kernel re-derives the surface frame from the same intermediate
// values (sphereSurface below).
#pragma once

#include "device_scene.h"
#include "intersect.cuh"
#include "libm_exact.cuh"

namespace iile {

__device__ __forceinline__ float nextUp(float v) {
    if (isinf(v) && v > 0.f) return v;
    if (v == -0.f) v = 0.f;
    uint32_t u = __float_as_uint(v);
    if (v >= 0) ++u; else --u;
    return __uint_as_float(u);
}
__device__ __forceinline__ float nextDown(float v) {
    if (isinf(v) && v < 0.f) return v;
    if (v == 0.f) v = -0.f;
    uint32_t u = __float_as_uint(v);
    if (v > 0) --u; else ++u;
    return __uint_as_float(u);
}

struct EF {  // efloat.h:48-70
    float v, lo, hi;
    __device__ EF() {}
    __device__ explicit EF(float v_, float err = 0.f) : v(v_) {
        if (err == 0.f) lo = hi = v_;
        else {
            lo = nextDown(v_ - err);
            hi = nextUp(v_ + err);
        }
    }
};
__device__ __forceinline__ EF efAdd(EF a, EF b) {
    EF r;
    r.v = a.v + b.v;
    r.lo = nextDown(a.lo + b.lo);
    r.hi = nextUp(a.hi + b.hi);
    return r;
}
__device__ __forceinline__ EF efSub(EF a, EF b) {
    EF r;
    r.v = a.v - b.v;
    r.lo = nextDown(a.lo - b.hi);
    r.hi = nextUp(a.hi - b.lo);
    return r;
}
__device__ __forceinline__ EF efMul(EF a, EF b) {
    EF r;
    r.v = a.v * b.v;
    float p0 = a.lo * b.lo, p1 = a.hi * b.lo, p2 = a.lo * b.hi, p3 = a.hi * b.hi;
    r.lo = nextDown(fminf(fminf(p0, p1), fminf(p2, p3)));
    r.hi = nextUp(fmaxf(fmaxf(p0, p1), fmaxf(p2, p3)));
    return r;
}
__device__ __forceinline__ EF efDiv(EF a, EF b) {
    EF r;
    r.v = a.v / b.v;
    if (b.lo < 0 && b.hi > 0) {
        r.lo = -INFINITY;
        r.hi = INFINITY;
    } else {
        float d0 = a.lo / b.lo, d1 = a.hi / b.lo, d2 = a.lo / b.hi, d3 = a.hi / b.hi;
        r.lo = nextDown(fminf(fminf(d0, d1), fminf(d2, d3)));
        r.hi = nextUp(fmaxf(fmaxf(d0, d1), fmaxf(d2, d3)));
    }
    return r;
}

// efloat.h:267-283
__device__ __forceinline__ bool efQuadratic(EF A, EF B, EF C, EF *t0, EF *t1) {
    double discrim = (double)B.v * (double)B.v - 4. * (double)A.v * (double)C.v;
    if (discrim < 0.) return false;
    double rootDiscrim = sqrt(discrim);
    EF fr((float)rootDiscrim, (float)((double)IILE_MACHINE_EPS * rootDiscrim));
    EF q;
    if (B.v < 0) q = efMul(EF(-.5f), efSub(B, fr));
    else q = efMul(EF(-.5f), efAdd(B, fr));
    *t0 = efDiv(q, A);
    *t1 = efDiv(C, q);
    if (t0->v > t1->v) {
        EF tmp = *t0;
        *t0 = *t1;
        *t1 = tmp;
    }
    return true;
}

struct SphereHit {
    float px, py, pz;  // refined object-space hit point
    float phi;
    float odx, ody, odz;  // object-space ray direction
    float u, v;           // parametric coordinates (sphere.cpp:100-103, disk.cpp:69-73, cylinder.cpp:108-110)
};

// Shared body of Sphere::Intersect / IntersectP up to the clipping tests.
// Disk and cylinder (disk.cpp:48-129, cylinder.cpp:48-200) in a function of their own: folded into
// sphereQuadric they raised the register demand of the call the traversal loop makes for every quadric
// leaf, and the closest-hit kernel paid 3 % in spills (profiles/r4f_ab_cast.txt).
static __device__ __noinline__ bool diskCylinderQuadric(const IileSphere &s, float rox, float roy, float roz,
                                                    float rdx, float rdy, float rdz, float tMax,
                                                    float *tOut, SphereHit *sh) {
    const float *m = s.worldToObject;
    const float g3 = gammaF(3);
    // point with error, transform.h:287-312
    float ox = m[0] * rox + m[1] * roy + m[2] * roz + m[3];
    float oy = m[4] * rox + m[5] * roy + m[6] * roz + m[7];
    float oz = m[8] * rox + m[9] * roy + m[10] * roz + m[11];
    float wp = m[12] * rox + m[13] * roy + m[14] * roz + m[15];
    float oex = g3 * (fabsf(m[0] * rox) + fabsf(m[1] * roy) + fabsf(m[2] * roz) + fabsf(m[3]));
    float oey = g3 * (fabsf(m[4] * rox) + fabsf(m[5] * roy) + fabsf(m[6] * roz) + fabsf(m[7]));
    float oez = g3 * (fabsf(m[8] * rox) + fabsf(m[9] * roy) + fabsf(m[10] * roz) + fabsf(m[11]));
    if (wp != 1) {
        float inv = 1.f / wp;
        ox *= inv; oy *= inv; oz *= inv;
    }
    // vector with error, transform.h:337-352
    float dex = g3 * (fabsf(m[0] * rdx) + fabsf(m[1] * rdy) + fabsf(m[2] * rdz));
    float dey = g3 * (fabsf(m[4] * rdx) + fabsf(m[5] * rdy) + fabsf(m[6] * rdz));
    float dez = g3 * (fabsf(m[8] * rdx) + fabsf(m[9] * rdy) + fabsf(m[10] * rdz));
    float dx = m[0] * rdx + m[1] * rdy + m[2] * rdz;
    float dy = m[4] * rdx + m[5] * rdy + m[6] * rdz;
    float dz = m[8] * rdx + m[9] * rdy + m[10] * rdz;
    // transform.h:370-381: push the origin to the edge of its error box
    float len2 = dx * dx + dy * dy + dz * dz;
    if (len2 > 0) {
        float dt = (fabsf(dx) * oex + fabsf(dy) * oey + fabsf(dz) * oez) / len2;
        ox += dx * dt; oy += dy * dt; oz += dz * dt;
    }
    if (s.flags & 4u) {  // Shape "disk": Disk::Intersect / IntersectP, disk.cpp:48-129 (height in zMin, inner radius in zMax)
        const float height = s.zMin, inner = s.zMax;
        if (dz == 0) return false;
        const float t = (height - oz) / dz;
        if (t <= 0 || t >= tMax) return false;
        const float hx = ox + dx * t, hy = oy + dy * t;
        const float dist2 = hx * hx + hy * hy;
        if (dist2 > s.radius * s.radius || dist2 < inner * inner) return false;
        float phi = atan2Glibc(hy, hx);  // std::atan2 of floats: the host's atan2f (libm_exact.cuh)
        if (phi < 0) phi += 2 * 3.14159265358979323846f;
        if (phi > s.phiMax) return false;
        *tOut = t;
        if (sh) {
            sh->px = hx; sh->py = hy; sh->pz = height;
            sh->phi = phi;
            sh->odx = dx; sh->ody = dy; sh->odz = dz;
            sh->u = phi / s.phiMax;
            sh->v = 1 - ((sqrtf(dist2) - inner) / (s.radius - inner));
        }
        return true;
    }
    if (s.flags & 8u) {  // Shape "cylinder": Cylinder::Intersect / IntersectP, cylinder.cpp:48-106,152-200
        EF eox(ox, oex), eoy(oy, oey);
        EF edx(dx, dex), edy(dy, dey);
        EF a = efAdd(efMul(edx, edx), efMul(edy, edy));
        EF b = efMul(EF(2.f), efAdd(efMul(edx, eox), efMul(edy, eoy)));
        EF c = efSub(efAdd(efMul(eox, eox), efMul(eoy, eoy)), efMul(EF(s.radius), EF(s.radius)));
        EF t0, t1;
        if (!efQuadratic(a, b, c, &t0, &t1)) return false;
        if (t0.hi > tMax || t1.lo <= 0) return false;
        EF tHit = t0;
        if (tHit.lo <= 0) {
            tHit = t1;
            if (tHit.hi > tMax) return false;
        }
        float px, py, pz, phi;
        auto refine = [&](float t) {
            px = ox + dx * t; py = oy + dy * t; pz = oz + dz * t;
            const float hitRad = sqrtf(px * px + py * py);
            px *= s.radius / hitRad;
            py *= s.radius / hitRad;
            phi = atan2Glibc(py, px);
            if (phi < 0) phi += 2 * 3.14159265358979323846f;
        };
        refine(tHit.v);
        if (pz < s.zMin || pz > s.zMax || phi > s.phiMax) {
            if (tHit.v == t1.v) return false;
            tHit = t1;
            if (t1.hi > tMax) return false;
            refine(tHit.v);
            if (pz < s.zMin || pz > s.zMax || phi > s.phiMax) return false;
        }
        *tOut = tHit.v;
        if (sh) {
            sh->px = px; sh->py = py; sh->pz = pz;
            sh->phi = phi;
            sh->odx = dx; sh->ody = dy; sh->odz = dz;
            sh->u = phi / s.phiMax;
            sh->v = (pz - s.zMin) / (s.zMax - s.zMin);
        }
        return true;
    }
    return false;
}

static __device__ __noinline__ bool sphereQuadric(const IileSphere &s, float rox, float roy, float roz,
                                              float rdx, float rdy, float rdz, float tMax,
                                              float *tOut, SphereHit *sh) {
    if (s.flags & 12u) return diskCylinderQuadric(s, rox, roy, roz, rdx, rdy, rdz, tMax, tOut, sh);
    const float *m = s.worldToObject;
    const float g3 = gammaF(3);
    // point with error, transform.h:287-312
    float ox = m[0] * rox + m[1] * roy + m[2] * roz + m[3];
    float oy = m[4] * rox + m[5] * roy + m[6] * roz + m[7];
    float oz = m[8] * rox + m[9] * roy + m[10] * roz + m[11];
    float wp = m[12] * rox + m[13] * roy + m[14] * roz + m[15];
    float oex = g3 * (fabsf(m[0] * rox) + fabsf(m[1] * roy) + fabsf(m[2] * roz) + fabsf(m[3]));
    float oey = g3 * (fabsf(m[4] * rox) + fabsf(m[5] * roy) + fabsf(m[6] * roz) + fabsf(m[7]));
    float oez = g3 * (fabsf(m[8] * rox) + fabsf(m[9] * roy) + fabsf(m[10] * roz) + fabsf(m[11]));
    if (wp != 1) {
        float inv = 1.f / wp;
        ox *= inv; oy *= inv; oz *= inv;
    }
    // vector with error, transform.h:337-352
    float dex = g3 * (fabsf(m[0] * rdx) + fabsf(m[1] * rdy) + fabsf(m[2] * rdz));
    float dey = g3 * (fabsf(m[4] * rdx) + fabsf(m[5] * rdy) + fabsf(m[6] * rdz));
    float dez = g3 * (fabsf(m[8] * rdx) + fabsf(m[9] * rdy) + fabsf(m[10] * rdz));
    float dx = m[0] * rdx + m[1] * rdy + m[2] * rdz;
    float dy = m[4] * rdx + m[5] * rdy + m[6] * rdz;
    float dz = m[8] * rdx + m[9] * rdy + m[10] * rdz;
    // transform.h:370-381: push the origin to the edge of its error box
    float len2 = dx * dx + dy * dy + dz * dz;
    if (len2 > 0) {
        float dt = (fabsf(dx) * oex + fabsf(dy) * oey + fabsf(dz) * oez) / len2;
        ox += dx * dt; oy += dy * dt; oz += dz * dt;
    }
    EF eox(ox, oex), eoy(oy, oey), eoz(oz, oez);
    EF edx(dx, dex), edy(dy, dey), edz(dz, dez);
    EF a = efAdd(efAdd(efMul(edx, edx), efMul(edy, edy)), efMul(edz, edz));
    EF b = efMul(EF(2.f), efAdd(efAdd(efMul(edx, eox), efMul(edy, eoy)), efMul(edz, eoz)));
    EF c = efSub(efAdd(efAdd(efMul(eox, eox), efMul(eoy, eoy)), efMul(eoz, eoz)),
                 efMul(EF(s.radius), EF(s.radius)));
    EF t0, t1;
    if (!efQuadratic(a, b, c, &t0, &t1)) return false;
    if (t0.hi > tMax || t1.lo <= 0) return false;
    EF tHit = t0;
    if (tHit.lo <= 0) {
        tHit = t1;
        if (tHit.hi > tMax) return false;
    }
    const float radius = s.radius;
    auto refine = [&](float t, float &px, float &py, float &pz, float &phi) {
        px = ox + dx * t; py = oy + dy * t; pz = oz + dz * t;
        float k = radius / sqrtf(px * px + py * py + pz * pz);
        px *= k; py *= k; pz *= k;
        if (px == 0 && py == 0) px = 1e-5f * radius;
        phi = atan2Glibc(py, px);  // std::atan2 of floats: the host's atan2f (libm_exact.cuh)
        if (phi < 0) phi += 2 * 3.14159265358979323846f;
    };
    float px, py, pz, phi;
    refine(tHit.v, px, py, pz, phi);
    if ((s.zMin > -radius && pz < s.zMin) || (s.zMax < radius && pz > s.zMax) || phi > s.phiMax) {
        if (tHit.v == t1.v) return false;
        if (t1.hi > tMax) return false;
        tHit = t1;
        refine(tHit.v, px, py, pz, phi);
        if ((s.zMin > -radius && pz < s.zMin) || (s.zMax < radius && pz > s.zMax) || phi > s.phiMax)
            return false;
    }
    *tOut = tHit.v;
    if (sh) {
        sh->px = px; sh->py = py; sh->pz = pz;
        sh->phi = phi;
        sh->odx = dx; sh->ody = dy; sh->odz = dz;
        float cz = pz / s.radius;
        cz = cz < -1.f ? -1.f : (cz > 1.f ? 1.f : cz);
        sh->u = phi / s.phiMax;
        sh->v = (acosGlibc(cz) - s.thetaMin) / (s.thetaMax - s.thetaMin);  // std::acos of a float: the host's acosf
    }
    return true;
}

// Hit test used by the traversal kernels; u,v are the parametric coordinates
// (sphere.cpp:100-103) stored in the hit record's b1,b2 lanes.
// Kept minimal: everything inlined here competes for the traversal loop's 72 registers (with the disk /
// cylinder (u, v) branches inlined the closest-hit kernel spilled 34 more bytes and lost 3 %), so the
// parametric coordinates are computed inside the out-of-line sphereQuadric.
__device__ __forceinline__ bool sphereTest(const IileSphere &s, const RayCtx &r, float tMax,
                                           float *tOut, float *uOut, float *vOut) {
    SphereHit sh;
    if (!sphereQuadric(s, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, tMax, tOut, &sh)) return false;
    *uOut = sh.u;
    *vOut = sh.v;
    return true;
}

}  // namespace iile
