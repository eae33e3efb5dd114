// Device-side shading substrate for the wavefront path integrator: surface
// reconstruction at a hit, the BSDF of the four in-scope materials, light
// sampling, ray spawning.  Each function restates the reference routine named
// in its comment with the same operation order (the TU is compiled with
// -fmad=false); the transcendental functions on the way (sinf / cosf / logf /
// atan2f / acosf) are restatements of the host glibc's algorithms (libm_exact.cuh),
// not CUDA's — every ray of a render carries the reference's bits
// (tests/test_rays_identical_gpu.py); films are sums of the same samples in another order.
#pragma once

#include "device_scene.h"
#include "intersect.cuh"
#include "pcg.cuh"
#include "sphere.cuh"
#include "libm_exact.cuh"

namespace iile {

#define IILE_PI 3.14159265358979323846f
#define IILE_INV_PI 0.31830988618379067154f
#define IILE_INV_2PI 0.15915494309189533577f
#define IILE_PI_OVER2 1.57079632679489661923f
#define IILE_PI_OVER4 0.78539816339744830961f
#define IILE_ONE_MINUS_EPS 0.99999994f
#define IILE_SHADOW_EPS 0.0001f

struct V3f {
    float x, y, z;
    __device__ V3f() {}
    __device__ V3f(float a, float b, float c) : x(a), y(b), z(c) {}
    __device__ float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
__device__ __forceinline__ V3f operator+(V3f a, V3f b) { return V3f(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3f operator-(V3f a, V3f b) { return V3f(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3f operator-(V3f a) { return V3f(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3f operator*(float s, V3f a) { return V3f(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3f operator*(V3f a, float s) { return V3f(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3f operator*(V3f a, V3f b) { return V3f(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3f operator/(V3f a, float f) {  // geometry.h: multiply by the reciprocal
    float inv = 1.f / f;
    return V3f(a.x * inv, a.y * inv, a.z * inv);
}
// CoefficientSpectrum::operator/(Float) divides each channel (spectrum.h:181-188)
__device__ __forceinline__ V3f sdiv(V3f a, float f) { return V3f(a.x / f, a.y / f, a.z / f); }
__device__ __forceinline__ float dot(V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float absDot(V3f a, V3f b) { return fabsf(dot(a, b)); }
__device__ __forceinline__ float lengthSq(V3f a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ float length(V3f a) { return sqrtf(lengthSq(a)); }
__device__ __forceinline__ V3f normalize(V3f a) { return a / length(a); }
__device__ __forceinline__ V3f cross(V3f a, V3f b) {  // geometry.h:957-963, evaluated in double
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V3f((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)),
               (float)((ax * by) - (ay * bx)));
}
__device__ __forceinline__ V3f faceforward(V3f n, V3f v) { return (dot(n, v) < 0.f) ? -n : n; }
__device__ __forceinline__ V3f absV(V3f a) { return V3f(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
__device__ __forceinline__ float distanceSq(V3f a, V3f b) { return lengthSq(a - b); }
__device__ __forceinline__ void coordinateSystem(V3f v1, V3f *v2, V3f *v3) {  // geometry.h:1020-1027
    if (fabsf(v1.x) > fabsf(v1.y)) *v2 = V3f(-v1.z, 0, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
    else *v2 = V3f(0, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
    *v3 = cross(v1, *v2);
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ float luminance(V3f c) {  // spectrum.h:463-466
    return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z;
}
__device__ __forceinline__ bool isBlack(V3f c) { return c.x == 0 && c.y == 0 && c.z == 0; }
__device__ __forceinline__ V3f clampSpectrum(V3f c) {  // Spectrum::Clamp(0, inf)
    return V3f(clampf(c.x, 0, INFINITY), clampf(c.y, 0, INFINITY), clampf(c.z, 0, INFINITY));
}
__device__ __forceinline__ V3f ld3(const float *p) { return V3f(p[0], p[1], p[2]); }

// OffsetRayOrigin, geometry.h:1440-1460
__device__ __forceinline__ V3f offsetRayOrigin(V3f p, V3f pError, V3f n, V3f w) {
    float d = dot(absV(n), pError);
    V3f offset = d * n;
    if (dot(w, n) < 0) offset = -offset;
    V3f po = p + offset;
    if (offset.x > 0) po.x = nextUp(po.x); else if (offset.x < 0) po.x = nextDown(po.x);
    if (offset.y > 0) po.y = nextUp(po.y); else if (offset.y < 0) po.y = nextDown(po.y);
    if (offset.z > 0) po.z = nextUp(po.z); else if (offset.z < 0) po.z = nextDown(po.z);
    return po;
}

// ---- libm on the device -------------------------------------------------------
// glibc's sinf / cosf / logf / acosf are correctly rounded in all but a few cases; CUDA's float versions are 1-2 ulp
// off.  IILE_LIBM_DOUBLE = 1 evaluates them in double and rounds once, which agrees with glibc nearly always: a
// bounce direction then carries the same bits as the reference's, instead of a last-ulp difference that every later
// vertex amplifies (A/B of parity and speed: profiles/r9_ab_libm_double.txt).
// IILE_LIBM_DOUBLE = 2 (default): sinf / cosf / logf / acosf / atan2f restated from glibc's own algorithms
// (libm_exact.cuh) — the bits of the host's libm, which is not correctly rounded itself.  With them every ray of a
// render is bit-identical to one the reference traces (tests/test_rays_identical_gpu.py).
#ifndef IILE_LIBM_DOUBLE
#define IILE_LIBM_DOUBLE 2
#endif
__device__ __forceinline__ void sincosLibm(float x, float *s, float *c) {
#if IILE_LIBM_DOUBLE == 2
    sincosGlibc(x, s, c);
#elif IILE_LIBM_DOUBLE
    double ds, dc;
    sincos((double)x, &ds, &dc);
    *s = (float)ds;
    *c = (float)dc;
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}
__device__ __forceinline__ float logLibm(float x) {
#if IILE_LIBM_DOUBLE == 2
    return logGlibc(x);
#elif IILE_LIBM_DOUBLE
    return (float)log((double)x);
#else
    return logf(x);
#endif
}
__device__ __forceinline__ float acosLibm(float x) {
#if IILE_LIBM_DOUBLE == 2
    return acosGlibc(x);
#elif IILE_LIBM_DOUBLE
    return (float)acos((double)x);
#else
    return acosf(x);
#endif
}

// ---- sampling helpers ---------------------------------------------------------
__device__ __forceinline__ void concentricSampleDisk(float u0, float u1, float *dx, float *dy) {
    // sampling.cpp:113-130
    float ox = 2.f * u0 - 1, oy = 2.f * u1 - 1;
    if (ox == 0 && oy == 0) { *dx = 0; *dy = 0; return; }
    float theta, r;
    if (fabsf(ox) > fabsf(oy)) {
        r = ox;
        theta = IILE_PI_OVER4 * (oy / ox);
    } else {
        r = oy;
        theta = IILE_PI_OVER2 - IILE_PI_OVER4 * (ox / oy);
    }
    float st, ct;
    sincosLibm(theta, &st, &ct);
    *dx = r * ct;
    *dy = r * st;
}
__device__ __forceinline__ V3f cosineSampleHemisphere(float u0, float u1) {  // sampling.h:159-163
    float dx, dy;
    concentricSampleDisk(u0, u1, &dx, &dy);
    float z = sqrtf(fmaxf(0.f, 1 - dx * dx - dy * dy));
    return V3f(dx, dy, z);
}
__device__ __forceinline__ float powerHeuristic(float fPdf, float gPdf) {  // sampling.h:171-174 (nf=ng=1)
    float f = 1 * fPdf, g = 1 * gPdf;
    return (f * f) / (f * f + g * g);
}

// ---- surface interaction --------------------------------------------------------
struct Surface {
    V3f p, pError;
    V3f n;        // geometric normal (oriented as the reference orients it)
    V3f ns;       // shading normal
    V3f ss, ts;   // BSDF frame: ss = Normalize(shading.dpdu), ts = Cross(ns, ss) (reflection.h:153-161)
    V3f wo;
    int material;
    int light;
};

// the part of a SurfaceInteraction that textures and Material::Bump read (texture.cuh)
struct TexGeom {
    float u, v;
    V3f dpdu, dpdv;      // geometric
    V3f shDpdu, shDpdv;  // shading.dpdu / shading.dpdv
    V3f shDndu, shDndv;  // shading.dndu / shading.dndv (the geometric dndu / dndv of a triangle are zero)
    float dudx, dudy, dvdx, dvdy;
};

// Triangle::Intersect's shading half (triangle.cpp:293-399) + BSDF frame.  The body is shared by the
// out-of-line triangleSurface (generic shade kernel) and the per-material kernels, which inline it.
// tg (textured materials only; a compile-time nullptr everywhere else): uv, dpdu / dpdv and the shading
// geometry SetShadingGeometry receives.
__device__ __forceinline__ void triangleSurfaceInl(const DeviceScene &sc, const IilePrim &pr, float b0,
                                                   float b1, float b2, V3f rayD, Surface *s, TexGeom *tg = nullptr) {
    const V3f p0 = ld3(sc.P + 3 * pr.v[0]), p1 = ld3(sc.P + 3 * pr.v[1]), p2 = ld3(sc.P + 3 * pr.v[2]);
    float uv[3][2] = {{0, 0}, {1, 0}, {1, 1}};  // triangle.h:98-108
    if (pr.flags & IILE_PRIM_HAS_UV)
        for (int k = 0; k < 3; ++k) {
            uv[k][0] = sc.UV[2 * pr.v[k]];
            uv[k][1] = sc.UV[2 * pr.v[k] + 1];
        }
    float duv02x = uv[0][0] - uv[2][0], duv02y = uv[0][1] - uv[2][1];
    float duv12x = uv[1][0] - uv[2][0], duv12y = uv[1][1] - uv[2][1];
    V3f dp02 = p0 - p2, dp12 = p1 - p2;
    float determinant = duv02x * duv12y - duv02y * duv12x;
    // triangle.cpp:302 compares the float with the DOUBLE literal 1e-8, which lies above 1e-8f: a determinant of
    // exactly 1e-8f is degenerate there (one comparison of immediates, no double arithmetic)
    bool degenerateUV = fabsf(determinant) <= 1e-8f;
    V3f dpdu(0, 0, 0), dpdv(0, 0, 0);
    if (!degenerateUV) {
        float invdet = 1 / determinant;
        dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
        dpdv = (-duv12x * dp02 + duv02x * dp12) * invdet;
    }
    if (degenerateUV || lengthSq(cross(dpdu, dpdv)) == 0)
        coordinateSystem(normalize(cross(p2 - p0, p1 - p0)), &dpdu, &dpdv);
    float xAbs = fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x);
    float yAbs = fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y);
    float zAbs = fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z);
    s->pError = gammaF(7) * V3f(xAbs, yAbs, zAbs);
    s->p = b0 * p0 + b1 * p1 + b2 * p2;
    s->wo = normalize(-rayD);  // Interaction ctor normalises wo (interaction.h:58-62)
    // SurfaceInteraction ctor: n = Normalize(Cross(dpdu,dpdv)) flipped by orientation; Triangle then
    // overrides n = shading.n = Normalize(Cross(dp02, dp12))
    const bool flip = (pr.flags & IILE_PRIM_FLIP_NORMAL) != 0;
    V3f n = normalize(cross(dp02, dp12));
    V3f ns = n;
    V3f shDpdu = dpdu;
    if (tg) {
        tg->u = b0 * uv[0][0] + b1 * uv[1][0] + b2 * uv[2][0];
        tg->v = b0 * uv[0][1] + b1 * uv[1][1] + b2 * uv[2][1];
        tg->dpdu = tg->shDpdu = dpdu;
        tg->dpdv = tg->shDpdv = dpdv;
        tg->shDndu = tg->shDndv = V3f(0, 0, 0);
    }
    if (pr.flags & IILE_PRIM_HAS_N) {
        const V3f n0 = ld3(sc.N + 3 * pr.v[0]), n1 = ld3(sc.N + 3 * pr.v[1]), n2 = ld3(sc.N + 3 * pr.v[2]);
        V3f nsI = b0 * n0 + b1 * n1 + b2 * n2;
        if (lengthSq(nsI) > 0) nsI = normalize(nsI);
        else nsI = n;
        V3f ss = normalize(dpdu);
        V3f ts = cross(ss, nsI);
        if (lengthSq(ts) > 0.f) {
            ts = normalize(ts);
            ss = cross(ts, nsI);
        } else
            coordinateSystem(nsI, &ss, &ts);
        if (tg) {  // triangle.cpp:371-390
            tg->shDpdu = ss;
            tg->shDpdv = ts;
            if (!degenerateUV) {
                const V3f dn1 = n0 - n2, dn2 = n1 - n2;
                const float invDet = 1 / determinant;
                tg->shDndu = (duv12y * dn1 - duv02y * dn2) * invDet;
                tg->shDndv = (-duv12x * dn1 + duv02x * dn2) * invDet;
            }
        }
        // SetShadingGeometry(ss, ts, ..., true), interaction.cpp:74-92
        V3f shn = normalize(cross(ss, ts));
        if (flip) shn = -shn;
        n = faceforward(n, shn);
        ns = shn;
        shDpdu = ss;
        // triangle.cpp:395-396: n = Faceforward(n, shading.n) again (idempotent)
        n = faceforward(n, ns);
    } else if (flip) {
        n = -n;
        ns = n;
    }
    s->n = n;
    s->ns = ns;
    s->ss = normalize(shDpdu);
    s->ts = cross(ns, s->ss);
    s->material = pr.material;
    s->light = pr.light;
}

static __device__ __noinline__ void triangleSurface(const DeviceScene &sc, const IilePrim &pr, float b0,
                                                float b1, float b2, V3f rayD, Surface *s) {
    triangleSurfaceInl(sc, pr, b0, b1, b2, rayD, s);
}

// Sphere::Intersect's parametric half + the object->world SurfaceInteraction transform
// (sphere.cpp:100-165, transform.cpp:257-291).
static __device__ __noinline__ void sphereSurface(const IileSphere &sp, const IilePrim &pr, const SphereHit &h,
                                              Surface *s) {
    const float radius = sp.radius;
    V3f pHit(h.px, h.py, h.pz);
    V3f dpdu(-sp.phiMax * pHit.y, sp.phiMax * pHit.x, 0);
    V3f dpdv, pErrObj;
    if (sp.flags & 4u) {  // Disk::Intersect's parametric half, disk.cpp:69-86 (pError = 0)
        const float rHit = sqrtf(pHit.x * pHit.x + pHit.y * pHit.y);
        dpdv = (V3f(pHit.x, pHit.y, 0.f) * (radius - sp.zMax)) / rHit;  // Vector3 / float: times the reciprocal
        pErrObj = V3f(0, 0, 0);
    } else if (sp.flags & 8u) {  // Cylinder::Intersect's parametric half, cylinder.cpp:112-139
        dpdv = V3f(0, 0, sp.zMax - sp.zMin);
        pErrObj = gammaF(3) * absV(V3f(pHit.x, pHit.y, 0));
    } else {
        float theta = acosLibm(clampf(pHit.z / radius, -1, 1));
        float zRadius = sqrtf(pHit.x * pHit.x + pHit.y * pHit.y);
        float invZRadius = 1 / zRadius;
        float cosPhi = pHit.x * invZRadius, sinPhi = pHit.y * invZRadius;
        float sinTheta, cosThetaUnused;
        sincosLibm(theta, &sinTheta, &cosThetaUnused);
        dpdv = (sp.thetaMax - sp.thetaMin) * V3f(pHit.z * cosPhi, pHit.z * sinPhi, -radius * sinTheta);
        pErrObj = gammaF(5) * absV(pHit);
    }
    // object-space SurfaceInteraction: n = Normalize(Cross(dpdu,dpdv)), flipped by orientation
    V3f nObj = normalize(cross(dpdu, dpdv));
    const bool flip = ((sp.flags & 1) != 0) ^ ((sp.flags & 2) != 0);
    if (flip) nObj = -nObj;
    const float *m = sp.objectToWorld, *mi = sp.worldToObject;
    // p with error, transform.h:314-335
    {
        float x = pHit.x, y = pHit.y, z = pHit.z;
        float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
        float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
        float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
        float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
        const float g3 = gammaF(3);
        s->pError.x = (g3 + 1.f) * (fabsf(m[0]) * pErrObj.x + fabsf(m[1]) * pErrObj.y + fabsf(m[2]) * pErrObj.z) +
                      g3 * (fabsf(m[0] * x) + fabsf(m[1] * y) + fabsf(m[2] * z) + fabsf(m[3]));
        s->pError.y = (g3 + 1.f) * (fabsf(m[4]) * pErrObj.x + fabsf(m[5]) * pErrObj.y + fabsf(m[6]) * pErrObj.z) +
                      g3 * (fabsf(m[4] * x) + fabsf(m[5] * y) + fabsf(m[6] * z) + fabsf(m[7]));
        s->pError.z = (g3 + 1.f) * (fabsf(m[8]) * pErrObj.x + fabsf(m[9]) * pErrObj.y + fabsf(m[10]) * pErrObj.z) +
                      g3 * (fabsf(m[8] * x) + fabsf(m[9] * y) + fabsf(m[10] * z) + fabsf(m[11]));
        s->p = (wp == 1.f) ? V3f(xp, yp, zp) : V3f(xp, yp, zp) / wp;
    }
    auto xfVec = [&](V3f v) {
        return V3f(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
                   m[8] * v.x + m[9] * v.y + m[10] * v.z);
    };
    auto xfNormal = [&](V3f v) {
        return V3f(mi[0] * v.x + mi[4] * v.y + mi[8] * v.z, mi[1] * v.x + mi[5] * v.y + mi[9] * v.z,
                   mi[2] * v.x + mi[6] * v.y + mi[10] * v.z);
    };
    V3f n = normalize(xfNormal(nObj));
    V3f ns = normalize(xfNormal(nObj));  // shading.n == n in object space
    ns = faceforward(ns, n);
    s->n = n;
    s->ns = ns;
    s->wo = normalize(xfVec(normalize(-V3f(h.odx, h.ody, h.odz))));
    s->ss = normalize(xfVec(dpdu));
    s->ts = cross(ns, s->ss);
    s->material = pr.material;
    s->light = pr.light;
}

// ---- BSDF of the in-scope materials ---------------------------------------------
enum : int { LOBE_LAMBERT = 0, LOBE_OREN_NAYAR = 1, LOBE_SPEC_REFL = 2, LOBE_FRESNEL_SPEC = 3, LOBE_MICROFACET = 4,
             LOBE_MICRO_TRANS = 5,     // MicrofacetTransmission (rough glass)
             LOBE_SPEC_REFL_DIEL = 6,  // SpecularReflection with FresnelDielectric (uber Kr)
             LOBE_SPEC_TRANS = 7,      // SpecularTransmission (uber opacity / Kt)
             LOBE_FRESNEL_BLEND = 8,   // FresnelBlend (substrate): R = Rd, T = Rs, a / b = alphax / alphay
             LOBE_LAMBERT_TRANS = 9 }; // LambertianTransmission (translucent): T
enum : int { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16,
             BSDF_ALL = 31 };

struct Lobe {
    int kind;
    int type;  // BxDFType bits
    V3f R;
    V3f T;      // FresnelSpecular
    float a, b;  // OrenNayar A,B | microfacet alphax, alphay | FresnelSpecular eta
    float etaA, etaB;  // FresnelDielectric(etaA, etaB) of the microfacet / specular lobes
};

constexpr int kMaxLobes = 5;  // uber: opacity transmission, diffuse, glossy, specular R, specular T

struct Bsdf {
    Lobe lobe[kMaxLobes];
    int n;
    float eta;
    V3f ns, ng, ss, ts;
    __device__ V3f toLocal(V3f v) const { return V3f(dot(v, ss), dot(v, ts), dot(v, ns)); }
    __device__ V3f toWorld(V3f v) const {
        return V3f(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z,
                   ss.z * v.x + ts.z * v.y + ns.z * v.z);
    }
    __device__ int numComponents(int flags) const {
        int c = 0;
        for (int i = 0; i < n; ++i)
            if ((lobe[i].type & flags) == lobe[i].type) ++c;
        return c;
    }
};

__device__ __forceinline__ float roughnessToAlpha(float roughness) {  // microfacet.h:123-128
    roughness = fmaxf(roughness, 1e-3f);
    float x = logLibm(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

// Material::ComputeScatteringFunctions for matte / plastic / mirror / glass
// (matte.cpp:45-62, plastic.cpp:45-70, mirror.cpp:45-56, glass.cpp:45-92 with allowMultipleLobes)
// Material::ComputeScatteringFunctions of substrate, translucent, uber, rough glass and of smooth glass
// without allowMultipleLobes: kept out of makeBsdf so that the four-material fast path stays compact.
static __device__ __noinline__ void makeBsdfRare(const IileMaterial &m, const Surface &s, Bsdf *b,
                                             bool allowMultipleLobes) {
    if (m.type == IILE_MAT_SUBSTRATE) {  // substrate.cpp:45-65
        const V3f d = clampSpectrum(ld3(m.Kd)), sp = clampSpectrum(ld3(m.Ks));
        if (!isBlack(d) || !isBlack(sp)) {
            float ru = m.uRoughness, rv = m.vRoughness;
            if (m.remapRoughness) {
                ru = roughnessToAlpha(ru);
                rv = roughnessToAlpha(rv);
            }
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_FRESNEL_BLEND;
            l.type = BSDF_REFLECTION | BSDF_GLOSSY;
            l.R = d;
            l.T = sp;
            l.a = ru;
            l.b = rv;
        }
    } else if (m.type == IILE_MAT_TRANSLUCENT) {  // translucent.cpp:45-82
        const float eta = 1.5f;
        b->eta = eta;
        const V3f r = clampSpectrum(ld3(m.Kr)), t = clampSpectrum(ld3(m.Kt));
        if (!(isBlack(r) && isBlack(t))) {
            const V3f kd = clampSpectrum(ld3(m.Kd));
            if (!isBlack(kd)) {
                if (!isBlack(r)) {
                    Lobe &l = b->lobe[b->n++];
                    l.kind = LOBE_LAMBERT;
                    l.type = BSDF_REFLECTION | BSDF_DIFFUSE;
                    l.R = r * kd;
                }
                if (!isBlack(t)) {
                    Lobe &l = b->lobe[b->n++];
                    l.kind = LOBE_LAMBERT_TRANS;
                    l.type = BSDF_TRANSMISSION | BSDF_DIFFUSE;
                    l.T = t * kd;
                }
            }
            const V3f ks = clampSpectrum(ld3(m.Ks));
            if (!isBlack(ks) && (!isBlack(r) || !isBlack(t))) {
                float rough = m.roughness;
                if (m.remapRoughness) rough = roughnessToAlpha(rough);
                if (!isBlack(r)) {
                    Lobe &l = b->lobe[b->n++];
                    l.kind = LOBE_MICROFACET;
                    l.type = BSDF_REFLECTION | BSDF_GLOSSY;
                    l.R = r * ks;
                    l.a = l.b = rough;
                    l.etaA = 1.f;
                    l.etaB = eta;
                }
                if (!isBlack(t)) {
                    Lobe &l = b->lobe[b->n++];
                    l.kind = LOBE_MICRO_TRANS;
                    l.type = BSDF_TRANSMISSION | BSDF_GLOSSY;
                    l.T = t * ks;
                    l.a = l.b = rough;
                    l.etaA = 1.f;
                    l.etaB = eta;
                }
            }
        }
    } else if (m.type == IILE_MAT_UBER) {  // uber.cpp:45-104
        const float e = m.eta;
        const V3f op = clampSpectrum(ld3(m.opacity));
        const V3f t = clampSpectrum(-op + V3f(1.f, 1.f, 1.f));
        if (!isBlack(t)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_SPEC_TRANS;
            l.type = BSDF_TRANSMISSION | BSDF_SPECULAR;
            l.T = t;
            l.etaA = l.etaB = 1.f;
        } else
            b->eta = e;
        const V3f kd = op * clampSpectrum(ld3(m.Kd));
        if (!isBlack(kd)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_LAMBERT;
            l.type = BSDF_REFLECTION | BSDF_DIFFUSE;
            l.R = kd;
        }
        const V3f ks = op * clampSpectrum(ld3(m.Ks));
        if (!isBlack(ks)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_MICROFACET;
            l.type = BSDF_REFLECTION | BSDF_GLOSSY;
            l.R = ks;
            float ru = m.uRoughness, rv = m.vRoughness;  // resolved by the loader (uber.cpp:71-79)
            if (m.remapRoughness) {
                ru = roughnessToAlpha(ru);
                rv = roughnessToAlpha(rv);
            }
            l.a = ru;
            l.b = rv;
            l.etaA = 1.f;
            l.etaB = e;
        }
        const V3f kr = op * clampSpectrum(ld3(m.Kr));
        if (!isBlack(kr)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_SPEC_REFL_DIEL;
            l.type = BSDF_REFLECTION | BSDF_SPECULAR;
            l.R = kr;
            l.etaA = 1.f;
            l.etaB = e;
        }
        const V3f kt = op * clampSpectrum(ld3(m.Kt));
        if (!isBlack(kt)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_SPEC_TRANS;
            l.type = BSDF_TRANSMISSION | BSDF_SPECULAR;
            l.T = kt;
            l.etaA = 1.f;
            l.etaB = e;
        }
    } else if (m.uRoughness != 0 || m.vRoughness != 0) {  // rough glass, glass.cpp:66-91
        V3f R = clampSpectrum(ld3(m.Ks)), T = clampSpectrum(ld3(m.Kt));
        b->eta = m.eta;
        if (!(isBlack(R) && isBlack(T))) {
            float ru = m.uRoughness, rv = m.vRoughness;
            if (m.remapRoughness) {
                ru = roughnessToAlpha(ru);
                rv = roughnessToAlpha(rv);
            }
            if (!isBlack(R)) {
                Lobe &l = b->lobe[b->n++];
                l.kind = LOBE_MICROFACET;
                l.type = BSDF_REFLECTION | BSDF_GLOSSY;
                l.R = R;
                l.a = ru;
                l.b = rv;
                l.etaA = 1.f;
                l.etaB = m.eta;
            }
            if (!isBlack(T)) {
                Lobe &l = b->lobe[b->n++];
                l.kind = LOBE_MICRO_TRANS;
                l.type = BSDF_TRANSMISSION | BSDF_GLOSSY;
                l.T = T;
                l.a = ru;
                l.b = rv;
                l.etaA = 1.f;
                l.etaB = m.eta;
            }
        }
    } else if (!allowMultipleLobes) {  // glass, smooth, one BxDF per lobe (glass.cpp:72-90 with isSpecular)
        V3f R = clampSpectrum(ld3(m.Ks)), T = clampSpectrum(ld3(m.Kt));
        b->eta = m.eta;
        if (!isBlack(R)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_SPEC_REFL_DIEL;
            l.type = BSDF_REFLECTION | BSDF_SPECULAR;
            l.R = R;
            l.etaA = 1.f;
            l.etaB = m.eta;
        }
        if (!isBlack(T)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_SPEC_TRANS;
            l.type = BSDF_TRANSMISSION | BSDF_SPECULAR;
            l.T = T;
            l.etaA = 1.f;
            l.etaB = m.eta;
        }
    }
}

// allowMultipleLobes: true for the path integrator (path.cpp:108), false for the direct-lighting
// integrators (interaction.h's default), where smooth glass is SpecularReflection + SpecularTransmission.
static __device__ __noinline__ void makeBsdf(const IileMaterial &m, const Surface &s, Bsdf *b,
                                         bool allowMultipleLobes = true) {
    b->n = 0;
    b->eta = 1;
    b->ns = s.ns;
    b->ng = s.n;
    b->ss = s.ss;
    b->ts = s.ts;
    if (m.type >= IILE_MAT_UBER ||
        (m.type == IILE_MAT_GLASS && (!allowMultipleLobes || m.uRoughness != 0 || m.vRoughness != 0))) {
        makeBsdfRare(m, s, b, allowMultipleLobes);
        return;
    }
    if (m.type == IILE_MAT_MATTE) {
        V3f r = clampSpectrum(ld3(m.Kd));
        float sig = clampf(m.sigma, 0, 90);
        if (!isBlack(r)) {
            Lobe &l = b->lobe[b->n++];
            l.R = r;
            l.type = BSDF_REFLECTION | BSDF_DIFFUSE;
            if (sig == 0) l.kind = LOBE_LAMBERT;
            else {  // OrenNayar ctor, reflection.h:412-419
                l.kind = LOBE_OREN_NAYAR;
                float sigma = (IILE_PI / 180.f) * sig;
                float sigma2 = sigma * sigma;
                l.a = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
                l.b = 0.45f * sigma2 / (sigma2 + 0.09f);
            }
        }
    } else if (m.type == IILE_MAT_PLASTIC) {
        V3f kd = clampSpectrum(ld3(m.Kd));
        if (!isBlack(kd)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_LAMBERT;
            l.type = BSDF_REFLECTION | BSDF_DIFFUSE;
            l.R = kd;
        }
        V3f ks = clampSpectrum(ld3(m.Ks));
        if (!isBlack(ks)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_MICROFACET;
            l.type = BSDF_REFLECTION | BSDF_GLOSSY;
            l.R = ks;
            float rough = m.roughness;
            if (m.remapRoughness) rough = roughnessToAlpha(rough);
            l.a = l.b = rough;
            l.etaA = 1.5f;  // plastic.cpp:63: FresnelDielectric(1.5f, 1.f)
            l.etaB = 1.f;
        }
    } else if (m.type == IILE_MAT_MIRROR) {
        V3f r = clampSpectrum(ld3(m.Ks));
        if (!isBlack(r)) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_SPEC_REFL;
            l.type = BSDF_REFLECTION | BSDF_SPECULAR;
            l.R = r;
        }
    } else {  // glass, smooth, allowMultipleLobes
        V3f R = clampSpectrum(ld3(m.Ks)), T = clampSpectrum(ld3(m.Kt));
        b->eta = m.eta;
        if (!(isBlack(R) && isBlack(T))) {
            Lobe &l = b->lobe[b->n++];
            l.kind = LOBE_FRESNEL_SPEC;
            l.type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
            l.R = R;
            l.T = T;
            l.a = m.eta;
        }
    }
}

// makeBsdf for one known material (the per-material shade kernels): the same lobes, built in line.
template <int MATERIAL>
__device__ __forceinline__ void makeBsdfFast(const IileMaterial &m, const Surface &s, Bsdf *b) {
    b->n = 0;
    b->eta = 1;
    b->ns = s.ns;
    b->ng = s.n;
    b->ss = s.ss;
    b->ts = s.ts;
    if (MATERIAL == IILE_MAT_MATTE) {
        V3f r = clampSpectrum(ld3(m.Kd));
        float sig = clampf(m.sigma, 0, 90);
        if (!isBlack(r)) {
            Lobe &l = b->lobe[b->n++];
            l.R = r;
            l.type = BSDF_REFLECTION | BSDF_DIFFUSE;
            if (sig == 0) l.kind = LOBE_LAMBERT;
            else {  // OrenNayar ctor, reflection.h:412-419
                l.kind = LOBE_OREN_NAYAR;
                float sigma = (IILE_PI / 180.f) * sig;
                float sigma2 = sigma * sigma;
                l.a = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
                l.b = 0.45f * sigma2 / (sigma2 + 0.09f);
            }
        }
    } else {  // IILE_MAT_PLASTIC; lobes are placed with compile-time indices so that the Bsdf stays in registers
        const V3f kd = clampSpectrum(ld3(m.Kd)), ks = clampSpectrum(ld3(m.Ks));
        Lobe diffuse{}, glossy{};
        diffuse.kind = LOBE_LAMBERT;
        diffuse.type = BSDF_REFLECTION | BSDF_DIFFUSE;
        diffuse.R = kd;
        glossy.kind = LOBE_MICROFACET;
        glossy.type = BSDF_REFLECTION | BSDF_GLOSSY;
        glossy.R = ks;
        float rough = m.roughness;
        if (m.remapRoughness) rough = roughnessToAlpha(rough);
        glossy.a = glossy.b = rough;
        glossy.etaA = 1.5f;  // plastic.cpp:63: FresnelDielectric(1.5f, 1.f)
        glossy.etaB = 1.f;
        const bool hasD = !isBlack(kd), hasG = !isBlack(ks);
        if (hasD && hasG) {
            b->lobe[0] = diffuse;
            b->lobe[1] = glossy;
            b->n = 2;
        } else if (hasD) {
            b->lobe[0] = diffuse;
            b->n = 1;
        } else if (hasG) {
            b->lobe[0] = glossy;
            b->n = 1;
        }
    }
}

// reflection.h:56-90 trigonometric helpers on local directions
__device__ __forceinline__ float cos2Theta(V3f w) { return w.z * w.z; }
__device__ __forceinline__ float sin2Theta(V3f w) { return fmaxf(0.f, 1.f - cos2Theta(w)); }
__device__ __forceinline__ float sinTheta(V3f w) { return sqrtf(sin2Theta(w)); }
__device__ __forceinline__ float tanTheta(V3f w) { return sinTheta(w) / w.z; }
__device__ __forceinline__ float tan2Theta(V3f w) { return sin2Theta(w) / cos2Theta(w); }
__device__ __forceinline__ float cosPhi(V3f w) {
    float st = sinTheta(w);
    return (st == 0) ? 1 : clampf(w.x / st, -1, 1);
}
__device__ __forceinline__ float sinPhi(V3f w) {
    float st = sinTheta(w);
    return (st == 0) ? 0 : clampf(w.y / st, -1, 1);
}
__device__ __forceinline__ bool sameHemisphere(V3f a, V3f b) { return a.z * b.z > 0; }

__device__ __forceinline__ float frDielectric(float cosThetaI, float etaI, float etaT) {  // reflection.cpp:47-68
    cosThetaI = clampf(cosThetaI, -1, 1);
    bool entering = cosThetaI > 0.f;
    if (!entering) {
        float t = etaI; etaI = etaT; etaT = t;
        cosThetaI = fabsf(cosThetaI);
    }
    float sinThetaI = sqrtf(fmaxf(0.f, 1 - cosThetaI * cosThetaI));
    float sinThetaT = etaI / etaT * sinThetaI;
    if (sinThetaT >= 1) return 1;
    float cosThetaT = sqrtf(fmaxf(0.f, 1 - sinThetaT * sinThetaT));
    float Rparl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
    float Rperp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
    return (Rparl * Rparl + Rperp * Rperp) / 2;
}

// TrowbridgeReitzDistribution(alphax, alphay), microfacet.cpp:155-184, 238-344
__device__ __forceinline__ float trD(V3f wh, float ax, float ay) {
    float t2 = tan2Theta(wh);
    if (isinf(t2)) return 0.f;
    float cos4 = cos2Theta(wh) * cos2Theta(wh);
    float cp = cosPhi(wh), spn = sinPhi(wh);
    float e = ((cp * cp) / (ax * ax) + (spn * spn) / (ay * ay)) * t2;
    return 1 / (IILE_PI * ax * ay * cos4 * (1 + e) * (1 + e));
}
__device__ __forceinline__ float trLambda(V3f w, float ax, float ay) {
    float absTan = fabsf(tanTheta(w));
    if (isinf(absTan)) return 0.f;
    float cp = cosPhi(w), spn = sinPhi(w);
    float a = sqrtf((cp * cp) * ax * ax + (spn * spn) * ay * ay);
    float a2t2 = (a * absTan) * (a * absTan);
    return (-1 + sqrtf(1.f + a2t2)) / 2;
}
__device__ __forceinline__ float trG1(V3f w, float ax, float ay) { return 1 / (1 + trLambda(w, ax, ay)); }
__device__ __forceinline__ float trG(V3f wo, V3f wi, float ax, float ay) {
    return 1 / (1 + trLambda(wo, ax, ay) + trLambda(wi, ax, ay));
}
__device__ __forceinline__ float trPdf(V3f wo, V3f wh, float ax, float ay) {  // sampleVisibleArea = true
    return trD(wh, ax, ay) * trG1(wo, ax, ay) * absDot(wo, wh) / fabsf(wo.z);
}
// Normal incidence (microfacet.cpp:241-247).  The reference writes this branch with unqualified sqrt / cos / sin and a
// double literal, so it is DOUBLE arithmetic there (microfacet.o imports `sincos`, not `sincosf`): phi = 2 pi U2 is a
// double product rounded to float, the slopes are r times the double cosine / sine, rounded once.  (sqrt in double
// and rounded is sqrtf.)  Out of line: the double sincos is code only low-roughness lobes reach.
static __device__ __noinline__ void trSample11Normal(float U1, float U2, float *slopeX, float *slopeY) {
    const float r = sqrtf(U1 / (1 - U1));
    const float phi = (float)(6.28318530718 * (double)U2);
    double sp, cp;
    sincos((double)phi, &sp, &cp);
    *slopeX = (float)((double)r * cp);
    *slopeY = (float)((double)r * sp);
}
__device__ __forceinline__ void trSample11(float cosTheta, float U1, float U2, float *slopeX, float *slopeY) {
    if (cosTheta > .9999f) {
        trSample11Normal(U1, U2, slopeX, slopeY);
        return;
    }
    float sinT = sqrtf(fmaxf(0.f, 1.f - cosTheta * cosTheta));
    float tanT = sinT / cosTheta;
    float a = 1 / tanT;
    float G1 = 2 / (1 + sqrtf(1.f + 1.f / (a * a)));
    float A = 2 * U1 / G1 - 1;
    float tmp = 1.f / (A * A - 1.f);
    if (tmp > 1e10f) tmp = 1e10f;
    float B = tanT;
    float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
    float sx1 = B * tmp - D, sx2 = B * tmp + D;
    *slopeX = (A < 0 || sx2 > 1.f / tanT) ? sx1 : sx2;
    float S;
    if (U2 > 0.5f) {
        S = 1.f;
        U2 = 2.f * (U2 - .5f);
    } else {
        S = -1.f;
        U2 = 2.f * (.5f - U2);
    }
    float z = (U2 * (U2 * (U2 * 0.27385f - 0.73369f) + 0.46341f)) /
              (U2 * (U2 * (U2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    *slopeY = S * z * sqrtf(1.f + *slopeX * *slopeX);
}
__device__ __forceinline__ V3f trSampleWh(V3f wo, float ax, float ay, float u0, float u1) {
    bool flip = wo.z < 0;
    V3f wi = flip ? -wo : wo;
    V3f ws = normalize(V3f(ax * wi.x, ay * wi.y, wi.z));
    float sx, sy;
    trSample11(ws.z, u0, u1, &sx, &sy);
    float tmp = cosPhi(ws) * sx - sinPhi(ws) * sy;
    sy = sinPhi(ws) * sx + cosPhi(ws) * sy;
    sx = tmp;
    sx = ax * sx;
    sy = ay * sy;
    V3f wh = normalize(V3f(-sx, -sy, 1.f));
    return flip ? -wh : wh;
}

// The lobes of rough glass, uber, substrate and translucent live in their own functions: the three
// dispatchers below stay as small as they were with four materials (the shade kernel is i-cache bound,
// profiles/r1g), and the code is fetched only by paths that meet those materials.
static __device__ __noinline__ V3f lobeFRare(const Lobe &l, V3f wo, V3f wi) {
    if (l.kind == LOBE_LAMBERT_TRANS) return l.T * IILE_INV_PI;  // reflection.cpp:188-191
    if (l.kind == LOBE_FRESNEL_BLEND) {  // FresnelBlend::f, reflection.cpp:285-298
        auto pow5 = [](float v) { return (v * v) * (v * v) * v; };
        const V3f one(1.f, 1.f, 1.f);
        V3f diffuse = (28.f / (23.f * IILE_PI)) * l.R * (one - l.T) * (1 - pow5(1 - .5f * fabsf(wi.z))) *
                      (1 - pow5(1 - .5f * fabsf(wo.z)));
        V3f wh = wi + wo;
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return V3f(0, 0, 0);
        wh = normalize(wh);
        const float q = trD(wh, l.a, l.b) / (4 * absDot(wi, wh) * fmaxf(fabsf(wi.z), fabsf(wo.z)));
        const V3f schlick = l.T + pow5(1 - dot(wi, wh)) * (one - l.T);
        return diffuse + q * schlick;
    }
    if (l.kind == LOBE_MICRO_TRANS) {  // MicrofacetTransmission::f, reflection.cpp:245-268 (TransportMode::Radiance)
        if (sameHemisphere(wo, wi)) return V3f(0, 0, 0);
        float cosO = wo.z, cosI = wi.z;
        if (cosI == 0 || cosO == 0) return V3f(0, 0, 0);
        float eta = wo.z > 0 ? (l.etaB / l.etaA) : (l.etaA / l.etaB);
        V3f wh = normalize(wo + wi * eta);
        if (wh.z < 0) wh = -wh;
        float F = frDielectric(dot(wo, wh), l.etaA, l.etaB);
        float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
        float factor = 1 / eta;
        V3f oneMinusF = V3f(1.f, 1.f, 1.f) - V3f(F, F, F);
        return oneMinusF * l.T *
               fabsf(trD(wh, l.a, l.b) * trG(wo, wi, l.a, l.b) * eta * eta * absDot(wi, wh) * absDot(wo, wh) * factor *
                     factor / (cosI * cosO * sqrtDenom * sqrtDenom));
    }
    return V3f(0, 0, 0);
}
static __device__ __noinline__ float lobePdfRare(const Lobe &l, V3f wo, V3f wi) {
    if (l.kind == LOBE_LAMBERT_TRANS) return !sameHemisphere(wo, wi) ? fabsf(wi.z) * IILE_INV_PI : 0;
    if (l.kind == LOBE_FRESNEL_BLEND) {  // FresnelBlend::Pdf, reflection.cpp:470-475
        if (!sameHemisphere(wo, wi)) return 0;
        V3f wh = normalize(wo + wi);
        const float pdfWh = trPdf(wo, wh, l.a, l.b);
        return .5f * (fabsf(wi.z) * IILE_INV_PI + pdfWh / (4 * dot(wo, wh)));
    }
    if (l.kind == LOBE_MICRO_TRANS) {  // MicrofacetTransmission::Pdf, reflection.cpp:448-461
        if (sameHemisphere(wo, wi)) return 0;
        float eta = wo.z > 0 ? (l.etaB / l.etaA) : (l.etaA / l.etaB);
        V3f wh = normalize(wo + wi * eta);
        float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
        float dwh_dwi = fabsf((eta * eta * dot(wi, wh)) / (sqrtDenom * sqrtDenom));
        return trPdf(wo, wh, l.a, l.b) * dwh_dwi;
    }
    return 0;
}
static __device__ __noinline__ V3f lobeSampleRare(const Lobe &l, V3f wo, float u0, float u1, V3f *wi, float *pdf,
                                              int *sampledType) {
    if (l.kind == LOBE_LAMBERT_TRANS) {  // LambertianTransmission::Sample_f, reflection.cpp:391-398
        *wi = cosineSampleHemisphere(u0, u1);
        if (wo.z > 0) wi->z *= -1;
        *pdf = lobePdfRare(l, wo, *wi);
        return lobeFRare(l, wo, *wi);
    }
    if (l.kind == LOBE_FRESNEL_BLEND) {  // FresnelBlend::Sample_f, reflection.cpp:450-468
        if (u0 < .5f) {
            u0 = fminf(2 * u0, IILE_ONE_MINUS_EPS);
            *wi = cosineSampleHemisphere(u0, u1);
            if (wo.z < 0) wi->z *= -1;
        } else {
            u0 = fminf(2 * (u0 - .5f), IILE_ONE_MINUS_EPS);
            V3f wh = trSampleWh(wo, l.a, l.b, u0, u1);
            *wi = -wo + 2 * dot(wo, wh) * wh;
            if (!sameHemisphere(wo, *wi)) return V3f(0, 0, 0);
        }
        *pdf = lobePdfRare(l, wo, *wi);
        return lobeFRare(l, wo, *wi);
    }
    if (l.kind == LOBE_MICRO_TRANS) {  // MicrofacetTransmission::Sample_f, reflection.cpp:436-446
        if (wo.z == 0) return V3f(0, 0, 0);
        V3f wh = trSampleWh(wo, l.a, l.b, u0, u1);
        float eta = wo.z > 0 ? (l.etaA / l.etaB) : (l.etaB / l.etaA);
        // Refract(wo, wh, eta, wi), reflection.h:96-108
        float cosThetaI = dot(wh, wo);
        float sin2ThetaI = fmaxf(0.f, 1 - cosThetaI * cosThetaI);
        float sin2ThetaT = eta * eta * sin2ThetaI;
        if (sin2ThetaT >= 1) return V3f(0, 0, 0);
        float cosThetaT = sqrtf(1 - sin2ThetaT);
        *wi = eta * -wo + (eta * cosThetaI - cosThetaT) * wh;
        *pdf = lobePdfRare(l, wo, *wi);
        return lobeFRare(l, wo, *wi);
    }
    if (l.kind == LOBE_SPEC_REFL_DIEL) {  // SpecularReflection::Sample_f, reflection.cpp:136-143
        *wi = V3f(-wo.x, -wo.y, wo.z);
        *pdf = 1;
        float F = frDielectric(wi->z, l.etaA, l.etaB);
        return sdiv(F * l.R, fabsf(wi->z));
    }
    if (l.kind == LOBE_SPEC_TRANS) {  // SpecularTransmission::Sample_f, reflection.cpp:151-168
        bool entering = wo.z > 0;
        float etaI = entering ? l.etaA : l.etaB, etaT = entering ? l.etaB : l.etaA;
        V3f n = faceforward(V3f(0, 0, 1), wo);
        float eta = etaI / etaT;
        float cosThetaI = dot(n, wo);
        float sin2ThetaI = fmaxf(0.f, 1 - cosThetaI * cosThetaI);
        float sin2ThetaT = eta * eta * sin2ThetaI;
        if (sin2ThetaT >= 1) return V3f(0, 0, 0);
        float cosThetaT = sqrtf(1 - sin2ThetaT);
        *wi = eta * -wo + (eta * cosThetaI - cosThetaT) * n;
        *pdf = 1;
        float F = frDielectric(wi->z, l.etaA, l.etaB);
        V3f ft = l.T * (V3f(1.f, 1.f, 1.f) - V3f(F, F, F));
        ft = ft * ((etaI * etaI) / (etaT * etaT));
        return sdiv(ft, fabsf(wi->z));
    }
    return V3f(0, 0, 0);
}

// KINDS: bit k set when lobe kind k can occur.  The generic entry points (lobeF, bsdfF, ... below) pass
// kAllLobes and stay out of line; the per-material shade kernels (wavefront.cu shadeFastKernel) pass the
// one or two kinds their material builds, and everything else folds away at compile time.  The
// arithmetic of a kind is the same expression in both.
constexpr int kAllLobes = 0x3ff;
constexpr int kMatteLobes = (1 << LOBE_LAMBERT) | (1 << LOBE_OREN_NAYAR);
constexpr int kPlasticLobes = (1 << LOBE_LAMBERT) | (1 << LOBE_MICROFACET);
constexpr int kRareLobes = kAllLobes & ~((1 << LOBE_MICRO_TRANS) - 1);
#define LOBE_CAN(kindBit) ((KINDS & (1 << (kindBit))) != 0)

// BxDF::f for the non-specular lobes (local frame)
template <int KINDS>
__device__ __forceinline__ V3f lobeFT(const Lobe &l, V3f wo, V3f wi) {
    if (LOBE_CAN(LOBE_LAMBERT) && l.kind == LOBE_LAMBERT) return l.R * IILE_INV_PI;
    if (LOBE_CAN(LOBE_OREN_NAYAR) && l.kind == LOBE_OREN_NAYAR) {  // reflection.cpp:197-219
        float sinI = sinTheta(wi), sinO = sinTheta(wo);
        float maxCos = 0;
        if (sinI > 1e-4f && sinO > 1e-4f) {
            float dCos = cosPhi(wi) * cosPhi(wo) + sinPhi(wi) * sinPhi(wo);
            maxCos = fmaxf(0.f, dCos);
        }
        float sinAlpha, tanBeta;
        if (fabsf(wi.z) > fabsf(wo.z)) {
            sinAlpha = sinO;
            tanBeta = sinI / fabsf(wi.z);
        } else {
            sinAlpha = sinI;
            tanBeta = sinO / fabsf(wo.z);
        }
        return l.R * IILE_INV_PI * (l.a + l.b * maxCos * sinAlpha * tanBeta);
    }
    if (LOBE_CAN(LOBE_MICROFACET) && l.kind == LOBE_MICROFACET) {  // reflection.cpp:226-236, FresnelDielectric(etaA, etaB)
        float cosO = fabsf(wo.z), cosI = fabsf(wi.z);
        V3f wh = wi + wo;
        if (cosI == 0 || cosO == 0) return V3f(0, 0, 0);
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return V3f(0, 0, 0);
        wh = normalize(wh);
        float F = frDielectric(dot(wi, wh), l.etaA, l.etaB);
        return sdiv(l.R * trD(wh, l.a, l.b) * trG(wo, wi, l.a, l.b) * F, 4 * cosI * cosO);
    }
    if ((KINDS & kRareLobes) && l.kind >= LOBE_MICRO_TRANS) return lobeFRare(l, wo, wi);  // rough glass / uber / substrate / translucent lobes
    return V3f(0, 0, 0);  // specular lobes: BxDF::f == 0
}
template <int KINDS>
__device__ __forceinline__ float lobePdfT(const Lobe &l, V3f wo, V3f wi) {
    if ((LOBE_CAN(LOBE_LAMBERT) && l.kind == LOBE_LAMBERT) || (LOBE_CAN(LOBE_OREN_NAYAR) && l.kind == LOBE_OREN_NAYAR))
        return sameHemisphere(wo, wi) ? fabsf(wi.z) * IILE_INV_PI : 0;
    if (LOBE_CAN(LOBE_MICROFACET) && l.kind == LOBE_MICROFACET) {
        if (!sameHemisphere(wo, wi)) return 0;
        V3f wh = normalize(wo + wi);
        return trPdf(wo, wh, l.a, l.b) / (4 * dot(wo, wh));
    }
    if ((KINDS & kRareLobes) && l.kind >= LOBE_MICRO_TRANS) return lobePdfRare(l, wo, wi);
    return 0;
}
// BxDF::Sample_f (local frame).  Returns f; sets wi, pdf, sampledType.
template <int KINDS>
__device__ __forceinline__ V3f lobeSampleT(const Lobe &l, V3f wo, float u0, float u1, V3f *wi, float *pdf,
                                           int *sampledType) {
    *sampledType = l.type;
    if ((LOBE_CAN(LOBE_LAMBERT) && l.kind == LOBE_LAMBERT) || (LOBE_CAN(LOBE_OREN_NAYAR) && l.kind == LOBE_OREN_NAYAR)) {  // reflection.cpp:378-389
        *wi = cosineSampleHemisphere(u0, u1);
        if (wo.z < 0) wi->z *= -1;
        *pdf = lobePdfT<KINDS>(l, wo, *wi);
        return lobeFT<KINDS>(l, wo, *wi);
    }
    if (LOBE_CAN(LOBE_MICROFACET) && l.kind == LOBE_MICROFACET) {  // reflection.cpp:405-417
        if (wo.z == 0) return V3f(0, 0, 0);
        V3f wh = trSampleWh(wo, l.a, l.b, u0, u1);
        *wi = -wo + 2 * dot(wo, wh) * wh;
        if (!sameHemisphere(wo, *wi)) return V3f(0, 0, 0);
        *pdf = trPdf(wo, wh, l.a, l.b) / (4 * dot(wo, wh));
        return lobeFT<KINDS>(l, wo, *wi);
    }
    if ((KINDS & kRareLobes) && l.kind >= LOBE_MICRO_TRANS) return lobeSampleRare(l, wo, u0, u1, wi, pdf, sampledType);
    if (!LOBE_CAN(LOBE_SPEC_REFL) && !LOBE_CAN(LOBE_FRESNEL_SPEC)) return V3f(0, 0, 0);
    if (l.kind == LOBE_SPEC_REFL) {  // reflection.cpp:136-143 with FresnelNoOp
        *wi = V3f(-wo.x, -wo.y, wo.z);
        *pdf = 1;
        return sdiv(1.f * l.R, fabsf(wi->z));
    }
    // FresnelSpecular, reflection.cpp:477-511 (etaA = 1, etaB = eta, TransportMode::Radiance)
    float F = frDielectric(wo.z, 1.f, l.a);
    if (u0 < F) {
        *wi = V3f(-wo.x, -wo.y, wo.z);
        *sampledType = BSDF_SPECULAR | BSDF_REFLECTION;
        *pdf = F;
        return sdiv(F * l.R, fabsf(wi->z));
    }
    bool entering = wo.z > 0;
    float etaI = entering ? 1.f : l.a, etaT = entering ? l.a : 1.f;
    // Refract(wo, Faceforward((0,0,1), wo), etaI/etaT, wi), reflection.h:96-108
    V3f n = faceforward(V3f(0, 0, 1), wo);
    float eta = etaI / etaT;
    float cosThetaI = dot(n, wo);
    float sin2ThetaI = fmaxf(0.f, 1 - cosThetaI * cosThetaI);
    float sin2ThetaT = eta * eta * sin2ThetaI;
    if (sin2ThetaT >= 1) return V3f(0, 0, 0);
    float cosThetaT = sqrtf(1 - sin2ThetaT);
    *wi = eta * -wo + (eta * cosThetaI - cosThetaT) * n;
    V3f ft = l.T * (1 - F);
    ft = ft * ((etaI * etaI) / (etaT * etaT));
    *sampledType = BSDF_SPECULAR | BSDF_TRANSMISSION;
    *pdf = 1 - F;
    return sdiv(ft, fabsf(wi->z));
}

static __device__ __noinline__ V3f lobeF(const Lobe &l, V3f wo, V3f wi) { return lobeFT<kAllLobes>(l, wo, wi); }
static __device__ __noinline__ float lobePdf(const Lobe &l, V3f wo, V3f wi) { return lobePdfT<kAllLobes>(l, wo, wi); }
static __device__ __noinline__ V3f lobeSample(const Lobe &l, V3f wo, float u0, float u1, V3f *wi, float *pdf,
                                              int *sampledType) {
    *sampledType = l.type;
    // the generic dispatcher calls the out-of-line f / pdf (as it always did): code size, not speed
    if (l.kind == LOBE_LAMBERT || l.kind == LOBE_OREN_NAYAR) {
        *wi = cosineSampleHemisphere(u0, u1);
        if (wo.z < 0) wi->z *= -1;
        *pdf = lobePdf(l, wo, *wi);
        return lobeF(l, wo, *wi);
    }
    if (l.kind == LOBE_MICROFACET) {
        if (wo.z == 0) return V3f(0, 0, 0);
        V3f wh = trSampleWh(wo, l.a, l.b, u0, u1);
        *wi = -wo + 2 * dot(wo, wh) * wh;
        if (!sameHemisphere(wo, *wi)) return V3f(0, 0, 0);
        *pdf = trPdf(wo, wh, l.a, l.b) / (4 * dot(wo, wh));
        return lobeF(l, wo, *wi);
    }
    constexpr int kSpecOnly = (1 << LOBE_SPEC_REFL) | (1 << LOBE_FRESNEL_SPEC) | kRareLobes;
    return lobeSampleT<kSpecOnly>(l, wo, u0, u1, wi, pdf, sampledType);
}

// BSDF::f, reflection.cpp:686-700.  MAXN: an upper bound of b.n known at compile time (the loops unroll
// and a small Bsdf lives in registers); NOINL: call the out-of-line lobe functions (the generic path).
template <int KINDS, int MAXN, bool NOINL>
__device__ __forceinline__ V3f bsdfFT(const Bsdf &b, V3f woW, V3f wiW, int flags) {
    V3f wi = b.toLocal(wiW), wo = b.toLocal(woW);
    if (wo.z == 0) return V3f(0, 0, 0);
    bool reflect = dot(wiW, b.ng) * dot(woW, b.ng) > 0;
    V3f f(0, 0, 0);
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        if (i >= b.n) break;
        const Lobe &l = b.lobe[i];
        if ((l.type & flags) == l.type &&
            ((reflect && (l.type & BSDF_REFLECTION)) || (!reflect && (l.type & BSDF_TRANSMISSION))))
            f = f + (NOINL ? lobeF(l, wo, wi) : lobeFT<KINDS>(l, wo, wi));
    }
    return f;
}
static __device__ __noinline__ V3f bsdfF(const Bsdf &b, V3f woW, V3f wiW, int flags) {
    return bsdfFT<kAllLobes, kMaxLobes, true>(b, woW, wiW, flags);
}
// BSDF::Pdf, reflection.cpp:786-801
template <int KINDS, int MAXN, bool NOINL>
__device__ __forceinline__ float bsdfPdfT(const Bsdf &b, V3f woW, V3f wiW, int flags) {
    if (b.n == 0) return 0.f;
    V3f wo = b.toLocal(woW), wi = b.toLocal(wiW);
    if (wo.z == 0) return 0.f;
    float pdf = 0.f;
    int matching = 0;
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        if (i >= b.n) break;
        if ((b.lobe[i].type & flags) == b.lobe[i].type) {
            ++matching;
            pdf += NOINL ? lobePdf(b.lobe[i], wo, wi) : lobePdfT<KINDS>(b.lobe[i], wo, wi);
        }
    }
    return matching > 0 ? pdf / matching : 0.f;
}
static __device__ __noinline__ float bsdfPdf(const Bsdf &b, V3f woW, V3f wiW, int flags) {
    return bsdfPdfT<kAllLobes, kMaxLobes, true>(b, woW, wiW, flags);
}
// BSDF::Sample_f, reflection.cpp:720-784
template <int KINDS, int MAXN, bool NOINL>
__device__ __forceinline__ V3f bsdfSampleT(const Bsdf &b, V3f woW, V3f *wiW, float u0, float u1, float *pdf,
                                           int flags, int *sampledType) {
    int matching = 0;
#pragma unroll
    for (int i = 0; i < MAXN; ++i)
        if (i < b.n && (b.lobe[i].type & flags) == b.lobe[i].type) ++matching;
    *sampledType = 0;
    if (matching == 0) {
        *pdf = 0;
        return V3f(0, 0, 0);
    }
    int comp = min((int)floorf(u0 * matching), matching - 1);
    int chosen = -1, count = comp;
#pragma unroll
    for (int i = 0; i < MAXN; ++i)
        if (chosen < 0 && i < b.n && (b.lobe[i].type & flags) == b.lobe[i].type && count-- == 0) chosen = i;
    float ur0 = fminf(u0 * matching - comp, IILE_ONE_MINUS_EPS);
    V3f wi, wo = b.toLocal(woW);
    *pdf = 0;
    if (wo.z == 0) return V3f(0, 0, 0);
    V3f f(0, 0, 0);
    bool spec = false;
    if (NOINL) {
        const Lobe &l = b.lobe[chosen];
        *sampledType = l.type;
        f = lobeSample(l, wo, ur0, u1, &wi, pdf, sampledType);
        spec = (l.type & BSDF_SPECULAR) != 0;
    } else {
#pragma unroll
        for (int i = 0; i < MAXN; ++i)  // the chosen lobe by a compile-time index: the Bsdf stays in registers
            if (i == chosen) {
                *sampledType = b.lobe[i].type;
                f = lobeSampleT<KINDS>(b.lobe[i], wo, ur0, u1, &wi, pdf, sampledType);
                spec = (b.lobe[i].type & BSDF_SPECULAR) != 0;
            }
    }
    if (*pdf == 0) {
        *sampledType = 0;
        return V3f(0, 0, 0);
    }
    *wiW = b.toWorld(wi);
    if (!spec && matching > 1) {
#pragma unroll
        for (int i = 0; i < MAXN; ++i)
            if (i < b.n && i != chosen && (b.lobe[i].type & flags) == b.lobe[i].type)
                *pdf += NOINL ? lobePdf(b.lobe[i], wo, wi) : lobePdfT<KINDS>(b.lobe[i], wo, wi);
    }
    if (matching > 1) *pdf /= matching;
    if (!spec && matching > 1) {
        bool reflect = dot(*wiW, b.ng) * dot(woW, b.ng) > 0;
        f = V3f(0, 0, 0);
#pragma unroll
        for (int i = 0; i < MAXN; ++i) {
            if (i >= b.n) break;
            const Lobe &k = b.lobe[i];
            if ((k.type & flags) == k.type &&
                ((reflect && (k.type & BSDF_REFLECTION)) || (!reflect && (k.type & BSDF_TRANSMISSION))))
                f = f + (NOINL ? lobeF(k, wo, wi) : lobeFT<KINDS>(k, wo, wi));
        }
    }
    return f;
}
static __device__ __noinline__ V3f bsdfSample(const Bsdf &b, V3f woW, V3f *wiW, float u0, float u1, float *pdf,
                                              int flags, int *sampledType) {
    return bsdfSampleT<kAllLobes, kMaxLobes, true>(b, woW, wiW, u0, u1, pdf, flags, sampledType);
}

// ---- lights ---------------------------------------------------------------------
struct LightSample {
    V3f Li, wi;
    float pdf;
    V3f pOrigin, pTarget;  // VisibilityTester end points after SpawnRayTo's offsets (interaction.h:73-78)
};

// The geometric normal the reference reports for a point on a triangle of an area light
// (Triangle::Sample, triangle.cpp:554-579, and Triangle::Intersect's orientation rule).
__device__ __forceinline__ V3f triangleLightNormal(const DeviceScene &sc, const IilePrim &pr, V3f p0, V3f p1,
                                                   V3f p2, float b0, float b1) {
    V3f n = normalize(cross(p1 - p0, p2 - p0));
    if (pr.flags & IILE_PRIM_HAS_N) {
        const V3f n0 = ld3(sc.N + 3 * pr.v[0]), n1 = ld3(sc.N + 3 * pr.v[1]), n2 = ld3(sc.N + 3 * pr.v[2]);
        V3f ns = b0 * n0 + b1 * n1 + (1 - b0 - b1) * n2;
        n = faceforward(n, ns);
    } else if (pr.flags & IILE_PRIM_FLIP_NORMAL)
        n = -n;
    return n;
}

__device__ __forceinline__ V3f xfPoint(const float *m, V3f p) {
    float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    return wp == 1 ? V3f(xp, yp, zp) : V3f(xp, yp, zp) / wp;
}

// ---- InfiniteAreaLight with a constant L (infinite.cpp; the 1x1 MIPMap's triangle filter, mipmap.h:264-274) ----
// glibc's sinf / cosf / acosf / atan2f are (nearly always) correctly rounded; CUDA's float versions
// are 1-2 ulp, so the trigonometry is evaluated in double and rounded once, as generateKernel does.
__device__ __forceinline__ V3f envLookup(const IileLight &L, float s0, float t0) {
    float s = s0 * 1 - 0.5f, t = t0 * 1 - 0.5f;
    float ds = s - (int)floorf(s), dt = t - (int)floorf(t);
    const V3f c = ld3(L.L);
    return (1 - ds) * (1 - dt) * c + (1 - ds) * dt * c + ds * (1 - dt) * c + ds * dt * c;
}
// InfiniteAreaLight::Le(ray), infinite.cpp:92-96
static __device__ __noinline__ V3f envLe(const IileLight &L, V3f d) {
    const float *m = L.worldToLight;
    V3f w = normalize(V3f(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z,
                          m[8] * d.x + m[9] * d.y + m[10] * d.z));
    float p = atan2Glibc(w.y, w.x);
    float phi = (p < 0) ? (p + 2 * IILE_PI) : p;
    float theta = acosLibm(clampf(w.z, -1, 1));
    return envLookup(L, phi * IILE_INV_2PI, theta * IILE_INV_PI);
}
// Distribution1D::SampleContinuous over two entries (sampling.h:73-92)
__device__ __forceinline__ float envSample1D(const float *func, const float *cdf, float funcInt, float u, float *pdf,
                                             int *off) {
    // FindInterval(3, cdf[i] <= u): clamp(first index with cdf > u, minus one, 0, 1)
    int offset = (cdf[1] <= u) ? 1 : 0;
    *off = offset;
    float du = u - cdf[offset];
    if ((cdf[offset + 1] - cdf[offset]) > 0) du /= (cdf[offset + 1] - cdf[offset]);
    *pdf = (funcInt > 0) ? func[offset] / funcInt : 0;
    return (offset + du) / 2;
}
// InfiniteAreaLight::Pdf_Li, infinite.cpp:124-133 (Distribution2D::Pdf, sampling.h:133-139)
static __device__ __noinline__ float envPdf(const IileLight &L, V3f w) {
    const float *m = L.worldToLight;
    V3f wi(m[0] * w.x + m[1] * w.y + m[2] * w.z, m[4] * w.x + m[5] * w.y + m[6] * w.z,
           m[8] * w.x + m[9] * w.y + m[10] * w.z);
    float theta = acosLibm(clampf(wi.z, -1, 1));
    float p = atan2Glibc(wi.y, wi.x);
    float phi = (p < 0) ? (p + 2 * IILE_PI) : p;
    float sinTheta, cosThetaUnused;
    sincosLibm(theta, &sinTheta, &cosThetaUnused);
    if (sinTheta == 0) return 0;
    int iu = min(max((int)(phi * IILE_INV_2PI * 2), 0), 1);
    int iv = min(max((int)(theta * IILE_INV_PI * 2), 0), 1);
    return (L.envFunc[iv][iu] / L.envMargFuncInt) / (2 * IILE_PI * IILE_PI * sinTheta);
}

// Light::Sample_Li for point, spot, distant, infinite and diffuse area lights (point.cpp:44-53, spot.cpp:53-75,
// distant.cpp:51-61, diffuse.cpp:68-81,
// shape.cpp:56-70, triangle.cpp:554-579, sphere.cpp:232-292).  ref: p, pError, n of the shading
// point (n = 0 for the light-grid build, lightdistrib.cpp:259-260).
static __device__ __noinline__ bool sampleLight(const DeviceScene &sc, const IileLight &L, V3f refP, V3f refPErr,
                                            V3f refN, float u0, float u1, LightSample *out) {
    if (L.type == IILE_LIGHT_POINT) {
        V3f pl = ld3(L.pos);
        out->wi = normalize(pl - refP);
        out->pdf = 1.f;
        out->Li = sdiv(ld3(L.L), distanceSq(pl, refP));
        out->pOrigin = offsetRayOrigin(refP, refPErr, refN, pl - refP);
        out->pTarget = pl;
        return true;
    }
    if (L.type == IILE_LIGHT_SPOT) {  // SpotLight::Sample_Li / Falloff, spot.cpp:53-75
        V3f pl = ld3(L.pos);
        out->wi = normalize(pl - refP);
        out->pdf = 1.f;
        const V3f w = -out->wi;
        const float *m = L.worldToLight;
        V3f wl = normalize(V3f(m[0] * w.x + m[1] * w.y + m[2] * w.z, m[4] * w.x + m[5] * w.y + m[6] * w.z,
                               m[8] * w.x + m[9] * w.y + m[10] * w.z));
        const float cosTheta = wl.z;
        float falloff;
        if (cosTheta < L.cosTotalWidth) falloff = 0;
        else if (cosTheta >= L.cosFalloffStart) falloff = 1;
        else {
            float delta = (cosTheta - L.cosTotalWidth) / (L.cosFalloffStart - L.cosTotalWidth);
            falloff = (delta * delta) * (delta * delta);
        }
        out->Li = sdiv(ld3(L.L) * falloff, distanceSq(pl, refP));
        out->pOrigin = offsetRayOrigin(refP, refPErr, refN, pl - refP);
        out->pTarget = pl;
        return true;
    }
    if (L.type == IILE_LIGHT_INFINITE) {  // InfiniteAreaLight::Sample_Li, infinite.cpp:98-122
        float pdfs[2];
        int v, unused;
        float d1 = envSample1D(L.envMargFunc, L.envMargCdf, L.envMargFuncInt, u1, &pdfs[1], &v);
        float d0 = envSample1D(L.envFunc[v], L.envCdf[v], L.envFuncInt[v], u0, &pdfs[0], &unused);
        float mapPdf = pdfs[0] * pdfs[1];
        out->pdf = 0;
        out->Li = V3f(0, 0, 0);
        if (mapPdf == 0) return true;
        float theta = d1 * IILE_PI, phi = d0 * 2 * IILE_PI;
        float sinTheta, cosTheta, sinPhi, cosPhi;
        sincosLibm(theta, &sinTheta, &cosTheta);
        sincosLibm(phi, &sinPhi, &cosPhi);
        const V3f wl(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
        const float *m = L.lightToWorld;
        out->wi = V3f(m[0] * wl.x + m[1] * wl.y + m[2] * wl.z, m[4] * wl.x + m[5] * wl.y + m[6] * wl.z,
                      m[8] * wl.x + m[9] * wl.y + m[10] * wl.z);
        out->pdf = mapPdf / (2 * IILE_PI * IILE_PI * sinTheta);
        if (sinTheta == 0) out->pdf = 0;
        const V3f pOutside = refP + out->wi * (2 * L.worldRadius);
        out->Li = envLookup(L, d0, d1);
        out->pOrigin = offsetRayOrigin(refP, refPErr, refN, pOutside - refP);
        out->pTarget = pOutside;
        return true;
    }
    if (L.type == IILE_LIGHT_DISTANT) {  // DistantLight::Sample_Li, distant.cpp:51-61
        const V3f wLight = ld3(L.pos);
        out->wi = wLight;
        out->pdf = 1.f;
        const V3f pOutside = refP + wLight * (2 * L.worldRadius);
        out->Li = ld3(L.L);
        out->pOrigin = offsetRayOrigin(refP, refPErr, refN, pOutside - refP);
        out->pTarget = pOutside;
        return true;
    }
    const IilePrim &pr = sc.prims[L.prim];
    V3f p, n, pErr;
    float pdf;
    if (pr.type == IILE_PRIM_TRIANGLE) {
        float su0 = sqrtf(u0);
        float b0 = 1 - su0, b1 = u1 * su0;
        const V3f p0 = ld3(sc.P + 3 * pr.v[0]), p1 = ld3(sc.P + 3 * pr.v[1]), p2 = ld3(sc.P + 3 * pr.v[2]);
        p = b0 * p0 + b1 * p1 + (1 - b0 - b1) * p2;
        n = triangleLightNormal(sc, pr, p0, p1, p2, b0, b1);
        V3f pAbs = absV(b0 * p0) + absV(b1 * p1) + absV((1 - b0 - b1) * p2);
        pErr = gammaF(6) * pAbs;
        pdf = 1 / L.area;
        // Shape::Sample(ref, u, pdf): area -> solid angle
        V3f wi = p - refP;
        if (lengthSq(wi) == 0) pdf = 0;
        else {
            wi = normalize(wi);
            pdf *= distanceSq(refP, p) / absDot(n, -wi);
            if (isinf(pdf)) pdf = 0.f;
        }
    } else {
        const IileSphere &sp = sc.spheres[pr.v[0]];
        const float radius = sp.radius;
        V3f pCenter = xfPoint(sp.objectToWorld, V3f(0, 0, 0));
        V3f pOrigin = offsetRayOrigin(refP, refPErr, refN, pCenter - refP);
        const bool rev = (sp.flags & 1) != 0;
        if (sp.flags & 4u) {
            // Disk::Sample(u, pdf), disk.cpp:136-146, then Shape::Sample(ref, u, pdf): area -> solid angle
            float dx, dy;
            concentricSampleDisk(u0, u1, &dx, &dy);
            V3f pObj(dx * radius, dy * radius, sp.zMin);
            const float *mi = sp.worldToObject;  // Normal3f(0,0,1) through the inverse transpose
            n = normalize(V3f(mi[8], mi[9], mi[10]));
            if (rev) n = -n;
            const float *m = sp.objectToWorld;
            const float g3 = gammaF(3);
            p = xfPoint(m, pObj);
            pErr.x = g3 * (fabsf(m[0] * pObj.x) + fabsf(m[1] * pObj.y) + fabsf(m[2] * pObj.z) + fabsf(m[3]));
            pErr.y = g3 * (fabsf(m[4] * pObj.x) + fabsf(m[5] * pObj.y) + fabsf(m[6] * pObj.z) + fabsf(m[7]));
            pErr.z = g3 * (fabsf(m[8] * pObj.x) + fabsf(m[9] * pObj.y) + fabsf(m[10] * pObj.z) + fabsf(m[11]));
            pdf = 1 / L.area;
            V3f wi = p - refP;
            if (lengthSq(wi) == 0) pdf = 0;
            else {
                wi = normalize(wi);
                pdf *= distanceSq(refP, p) / absDot(n, -wi);
                if (isinf(pdf)) pdf = 0.f;
            }
        } else if (sp.flags & 8u) {
            // Cylinder::Sample(u, pdf), cylinder.cpp:205-220, then Shape::Sample(ref, u, pdf)
            const float z = (1 - u0) * sp.zMin + u0 * sp.zMax;
            const float phi = u1 * sp.phiMax;
            float sphi, cphi;
            sincosLibm(phi, &sphi, &cphi);
            V3f pObj(radius * cphi, radius * sphi, z);
            const float *mi = sp.worldToObject;
            n = normalize(V3f(mi[0] * pObj.x + mi[4] * pObj.y + mi[8] * 0.f,
                              mi[1] * pObj.x + mi[5] * pObj.y + mi[9] * 0.f,
                              mi[2] * pObj.x + mi[6] * pObj.y + mi[10] * 0.f));
            if (rev) n = -n;
            const float hitRad = sqrtf(pObj.x * pObj.x + pObj.y * pObj.y);
            pObj.x *= radius / hitRad;
            pObj.y *= radius / hitRad;
            V3f pObjErr = gammaF(3) * absV(V3f(pObj.x, pObj.y, 0));
            const float *m = sp.objectToWorld;
            const float g3 = gammaF(3);
            p = xfPoint(m, pObj);
            pErr.x = (g3 + 1.f) * (fabsf(m[0]) * pObjErr.x + fabsf(m[1]) * pObjErr.y + fabsf(m[2]) * pObjErr.z) +
                     g3 * (fabsf(m[0] * pObj.x) + fabsf(m[1] * pObj.y) + fabsf(m[2] * pObj.z) + fabsf(m[3]));
            pErr.y = (g3 + 1.f) * (fabsf(m[4]) * pObjErr.x + fabsf(m[5]) * pObjErr.y + fabsf(m[6]) * pObjErr.z) +
                     g3 * (fabsf(m[4] * pObj.x) + fabsf(m[5] * pObj.y) + fabsf(m[6] * pObj.z) + fabsf(m[7]));
            pErr.z = (g3 + 1.f) * (fabsf(m[8]) * pObjErr.x + fabsf(m[9]) * pObjErr.y + fabsf(m[10]) * pObjErr.z) +
                     g3 * (fabsf(m[8] * pObj.x) + fabsf(m[9] * pObj.y) + fabsf(m[10] * pObj.z) + fabsf(m[11]));
            pdf = 1 / L.area;
            V3f wi = p - refP;
            if (lengthSq(wi) == 0) pdf = 0;
            else {
                wi = normalize(wi);
                pdf *= distanceSq(refP, p) / absDot(n, -wi);
                if (isinf(pdf)) pdf = 0.f;
            }
        } else if (distanceSq(pOrigin, pCenter) <= radius * radius) {
            // Sphere::Sample(u, pdf), sphere.cpp:219-230, then area -> solid angle
            float z = 1 - 2 * u0;
            float r = sqrtf(fmaxf(0.f, 1.f - z * z));
            float phi = 2 * IILE_PI * u1;
            float sphi, cphi;
            sincosLibm(phi, &sphi, &cphi);
            V3f pObj = V3f(0, 0, 0) + radius * V3f(r * cphi, r * sphi, z);
            const float *mi = sp.worldToObject;
            n = normalize(V3f(mi[0] * pObj.x + mi[4] * pObj.y + mi[8] * pObj.z,
                              mi[1] * pObj.x + mi[5] * pObj.y + mi[9] * pObj.z,
                              mi[2] * pObj.x + mi[6] * pObj.y + mi[10] * pObj.z));
            if (rev) n = -n;
            pObj = pObj * (radius / length(pObj));
            V3f pObjErr = gammaF(5) * absV(pObj);
            const float *m = sp.objectToWorld;
            const float g3 = gammaF(3);
            p = xfPoint(m, pObj);
            pErr.x = (g3 + 1.f) * (fabsf(m[0]) * pObjErr.x + fabsf(m[1]) * pObjErr.y + fabsf(m[2]) * pObjErr.z) +
                     g3 * (fabsf(m[0] * pObj.x) + fabsf(m[1] * pObj.y) + fabsf(m[2] * pObj.z) + fabsf(m[3]));
            pErr.y = (g3 + 1.f) * (fabsf(m[4]) * pObjErr.x + fabsf(m[5]) * pObjErr.y + fabsf(m[6]) * pObjErr.z) +
                     g3 * (fabsf(m[4] * pObj.x) + fabsf(m[5] * pObj.y) + fabsf(m[6] * pObj.z) + fabsf(m[7]));
            pErr.z = (g3 + 1.f) * (fabsf(m[8]) * pObjErr.x + fabsf(m[9]) * pObjErr.y + fabsf(m[10]) * pObjErr.z) +
                     g3 * (fabsf(m[8] * pObj.x) + fabsf(m[9] * pObj.y) + fabsf(m[10] * pObj.z) + fabsf(m[11]));
            pdf = 1 / L.area;
            V3f wi = p - refP;
            if (lengthSq(wi) == 0) pdf = 0;
            else {
                wi = normalize(wi);
                pdf *= distanceSq(refP, p) / absDot(n, -wi);
            }
            if (isinf(pdf)) pdf = 0.f;
        } else {
            // cone sampling, sphere.cpp:252-290
            V3f wc = normalize(pCenter - refP);
            V3f wcX, wcY;
            coordinateSystem(wc, &wcX, &wcY);
            float sinThetaMax2 = radius * radius / distanceSq(refP, pCenter);
            float cosThetaMax = sqrtf(fmaxf(0.f, 1 - sinThetaMax2));
            float cosTheta = (1 - u0) + u0 * cosThetaMax;
            float sinT = sqrtf(fmaxf(0.f, 1 - cosTheta * cosTheta));
            float phi = u1 * 2 * IILE_PI;
            float dc = length(refP - pCenter);
            float ds = dc * cosTheta - sqrtf(fmaxf(0.f, radius * radius - dc * dc * sinT * sinT));
            float cosAlpha = (dc * dc + radius * radius - ds * ds) / (2 * dc * radius);
            float sinAlpha = sqrtf(fmaxf(0.f, 1 - cosAlpha * cosAlpha));
            float sphi, cphi;
            sincosLibm(phi, &sphi, &cphi);
            V3f nW = sinAlpha * cphi * (-wcX) + sinAlpha * sphi * (-wcY) + cosAlpha * (-wc);
            p = pCenter + radius * nW;
            pErr = gammaF(5) * absV(p);
            n = nW;
            if (rev) n = -n;
            pdf = 1 / (2 * IILE_PI * (1 - cosThetaMax));
        }
    }
    // DiffuseAreaLight::Sample_Li, diffuse.cpp:68-81
    if (pdf == 0 || lengthSq(p - refP) == 0) {
        out->pdf = 0;
        out->Li = V3f(0, 0, 0);
        return false;
    }
    out->wi = normalize(p - refP);
    out->pdf = pdf;
    out->Li = (IILE_LIGHT_TWOSIDED(L) || dot(n, -out->wi) > 0) ? ld3(L.L) : V3f(0, 0, 0);
    // VisibilityTester(ref, pShape) -> SpawnRayTo(it): the target is offset towards the origin
    out->pOrigin = offsetRayOrigin(refP, refPErr, refN, p - refP);
    out->pTarget = offsetRayOrigin(p, pErr, n, out->pOrigin - p);
    return true;
}

// sampleLight for a scene whose lights are all point lights or diffuse area lights on triangles (what
// the per-material shade kernels of such scenes inline): the same two branches, nothing else.
__device__ __forceinline__ bool sampleLightSimple(const DeviceScene &sc, const IileLight &L, V3f refP, V3f refPErr,
                                                  V3f refN, float u0, float u1, LightSample *out) {
    if (L.type == IILE_LIGHT_POINT) {
        V3f pl = ld3(L.pos);
        out->wi = normalize(pl - refP);
        out->pdf = 1.f;
        out->Li = sdiv(ld3(L.L), distanceSq(pl, refP));
        out->pOrigin = offsetRayOrigin(refP, refPErr, refN, pl - refP);
        out->pTarget = pl;
        return true;
    }
    const IilePrim &pr = sc.prims[L.prim];
    V3f p, n, pErr;
    float pdf;
    {
        float su0 = sqrtf(u0);
        float b0 = 1 - su0, b1 = u1 * su0;
        const V3f p0 = ld3(sc.P + 3 * pr.v[0]), p1 = ld3(sc.P + 3 * pr.v[1]), p2 = ld3(sc.P + 3 * pr.v[2]);
        p = b0 * p0 + b1 * p1 + (1 - b0 - b1) * p2;
        n = triangleLightNormal(sc, pr, p0, p1, p2, b0, b1);
        V3f pAbs = absV(b0 * p0) + absV(b1 * p1) + absV((1 - b0 - b1) * p2);
        pErr = gammaF(6) * pAbs;
        pdf = 1 / L.area;
        V3f wi = p - refP;
        if (lengthSq(wi) == 0) pdf = 0;
        else {
            wi = normalize(wi);
            pdf *= distanceSq(refP, p) / absDot(n, -wi);
            if (isinf(pdf)) pdf = 0.f;
        }
    }
    if (pdf == 0 || lengthSq(p - refP) == 0) {
        out->pdf = 0;
        out->Li = V3f(0, 0, 0);
        return false;
    }
    out->wi = normalize(p - refP);
    out->pdf = pdf;
    out->Li = (IILE_LIGHT_TWOSIDED(L) || dot(n, -out->wi) > 0) ? ld3(L.L) : V3f(0, 0, 0);
    out->pOrigin = offsetRayOrigin(refP, refPErr, refN, p - refP);
    out->pTarget = offsetRayOrigin(p, pErr, n, out->pOrigin - p);
    return true;
}

// areaLightPdf for a diffuse area light on a triangle, in line
__device__ __forceinline__ float areaLightPdfTriangle(const DeviceScene &sc, const IileLight &L, V3f refP, V3f spawnO,
                                                      V3f wi) {
    const IilePrim &pr = sc.prims[L.prim];
    RayCtx r;
    initRay(r, spawnO.x, spawnO.y, spawnO.z, wi.x, wi.y, wi.z, INFINITY);
    const V3f p0 = ld3(sc.P + 3 * pr.v[0]), p1 = ld3(sc.P + 3 * pr.v[1]), p2 = ld3(sc.P + 3 * pr.v[2]);
    float t, b0, b1, b2;
    if (!triangleTest(r, INFINITY, make_float3(p0.x, p0.y, p0.z), make_float3(p1.x, p1.y, p1.z),
                      make_float3(p2.x, p2.y, p2.z), &t, &b0, &b1, &b2))
        return 0;
    Surface s;
    triangleSurfaceInl(sc, pr, b0, b1, b2, wi, &s);
    float pdf = distanceSq(refP, s.p) / (absDot(s.n, -wi) * L.area);
    if (isinf(pdf)) pdf = 0.f;
    return pdf;
}

// Light::Pdf_Li(ref, wi) for a diffuse area light = Shape::Pdf (shape.cpp:72-87, sphere.cpp:294-306).
// spawnO: ref.SpawnRay(wi).o
static __device__ __noinline__ float areaLightPdf(const DeviceScene &sc, const IileLight &L, V3f refP, V3f refPErr,
                                              V3f refN, V3f spawnO, V3f wi) {
    const IilePrim &pr = sc.prims[L.prim];
    RayCtx r;
    initRay(r, spawnO.x, spawnO.y, spawnO.z, wi.x, wi.y, wi.z, INFINITY);
    V3f pHit, nHit;
    if (pr.type == IILE_PRIM_TRIANGLE) {
        const V3f p0 = ld3(sc.P + 3 * pr.v[0]), p1 = ld3(sc.P + 3 * pr.v[1]), p2 = ld3(sc.P + 3 * pr.v[2]);
        float t, b0, b1, b2;
        if (!triangleTest(r, INFINITY, make_float3(p0.x, p0.y, p0.z), make_float3(p1.x, p1.y, p1.z),
                          make_float3(p2.x, p2.y, p2.z), &t, &b0, &b1, &b2))
            return 0;
        Surface s;
        triangleSurface(sc, pr, b0, b1, b2, wi, &s);
        pHit = s.p;
        nHit = s.n;
    } else {
        const IileSphere &sp = sc.spheres[pr.v[0]];
        V3f pCenter = xfPoint(sp.objectToWorld, V3f(0, 0, 0));
        V3f pOrigin = offsetRayOrigin(refP, refPErr, refN, pCenter - refP);
        if (!(sp.flags & 12u) && !(distanceSq(pOrigin, pCenter) <= sp.radius * sp.radius)) {
            float sinThetaMax2 = sp.radius * sp.radius / distanceSq(refP, pCenter);
            float cosThetaMax = sqrtf(fmaxf(0.f, 1 - sinThetaMax2));
            return 1 / (2 * IILE_PI * (1 - cosThetaMax));
        }
        float t;
        SphereHit sh;
        if (!sphereQuadric(sp, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, INFINITY, &t, &sh)) return 0;
        Surface s;
        sphereSurface(sp, pr, sh, &s);
        pHit = s.p;
        nHit = s.n;
    }
    float pdf = distanceSq(refP, pHit) / (absDot(nHit, -wi) * L.area);
    if (isinf(pdf)) pdf = 0.f;
    return pdf;
}

#ifndef IILE_DISCRETE_LINEAR
#define IILE_DISCRETE_LINEAR 1
#endif
// Distribution1D::SampleDiscrete on a stored table {cdf[n+1], func[n], funcInt} (sampling.h:90-100)
__device__ __forceinline__ int sampleDiscrete(const float *table, int n, float u, float *pdf) {
    const float *cdf = table, *func = table + n + 1;
    const float funcInt = table[2 * n + 1];
    int size = n + 1, first = 0, len = size;
#if IILE_DISCRETE_LINEAR
    // FindInterval's binary search (pbrt.h:383-397) ends at the first entry with cdf > u; the cdf is
    // non-decreasing, so that index is the number of entries <= u — counted here with independent loads
    // instead of five dependent ones (ncu: 4.3 % of the matte shade kernel's stall samples sat on the search's
    // load, profiles/r8g_shade_cfg3_*).  Same answer for every u.
    if (size <= 64) {
        len = 0;
#pragma unroll 4
        for (int i = 0; i < size; ++i) first += cdf[i] <= u ? 1 : 0;
    }
#endif
    while (len > 0) {
        int half = len >> 1, middle = first + half;
        if (cdf[middle] <= u) {
            first = middle + 1;
            len -= half + 1;
        } else
            len = half;
    }
    int offset = min(max(first - 1, 0), size - 2);
    *pdf = (funcInt > 0) ? func[offset] / (funcInt * n) : 0;
    return offset;
}

// LightDistribution::Lookup, lightdistrib.cpp:69-81, 134-147
__device__ __forceinline__ const float *lightTableAt(const DeviceScene &sc, V3f p) {
    if (sc.lightStrategy != IILE_LIGHTSTRATEGY_SPATIAL) return sc.lightCdf;
    int vi[3];
    for (int a = 0; a < 3; ++a) {
        float lo = sc.worldMin[a], hi = sc.worldMax[a];
        float o = p[a] - lo;
        if (hi > lo) o /= hi - lo;
        int v = (int)(o * sc.voxels[a]);
        vi[a] = min(max(v, 0), sc.voxels[a] - 1);
    }
    size_t voxel = ((size_t)vi[2] * sc.voxels[1] + vi[1]) * sc.voxels[0] + vi[0];
    return sc.lightCdf + voxel * (2 * (size_t)sc.nLights + 2);
}

}  // namespace iile
