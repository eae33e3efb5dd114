// Textured materials on the device (SURVEY §8 a14 + f2): ray differentials, texture evaluation, bump mapping.
//
//   camera differentials   PerspectiveCamera / OrthographicCamera::GenerateRayDifferential (perspective.cpp:100-149,
//                          orthographic.cpp:69-118), the base class's finite differences for the environment
//                          camera (camera.cpp:59-94),
//                          RayDifferential::ScaleDifferentials (geometry.h:908-913, integrator.cpp:278-279)
//   ComputeDifferentials   interaction.cpp:103-149: dpdx, dpdy, dudx, dvdx, dudy, dvdy at a camera-ray hit;
//                          every later vertex of a path has none (path.cpp:157 spawns a plain Ray)
//   textures               ConstantTexture, ScaleTexture (scale.h:56-58), ImageTexture::Evaluate with
//                          UVMapping2D::Map (imagemap.h:82-89, texture.cpp:93-99) and MIPMap::Lookup —
//                          trilinear and EWA (mipmap.h:239-369) — over the pyramid the host built
//                          (host/texture_host.cpp, bit-identical to the reference's)
//   bump mapping           Material::Bump (material.cpp:46-83) + SetShadingGeometry (interaction.cpp:73-92)
//
// Only materials with a texture set come here (they are classed kClassOther, so the per-material shade
// kernels never see them): resolveTextured() evaluates the textures of the material's slots into a copy of
// the material record — which makeBsdf then reads exactly as it reads a constant material — and applies
// the bump map to the surface's shading frame.
#pragma once

#include "camera.cuh"
#include "intersect.cuh"
#include "shading.cuh"
#include "wavefront.h"

namespace iile {

__device__ __forceinline__ int texMod(int a, int b) {  // pbrt.h Mod
    const int r = a - (a / b) * b;
    return r < 0 ? r + b : r;
}

struct MipLevel {
    const float *texels;
    int w, h;
};
__device__ __forceinline__ MipLevel mipLevel(const DeviceScene &sc, const IileImage &im, int level) {
    size_t at = im.texelOffset;
    int w = im.width, h = im.height;
    for (int l = 0; l < level; ++l) {
        at += (size_t)w * h * im.channels;
        w = max(1, w >> 1);
        h = max(1, h >> 1);
    }
    return MipLevel{sc.texels + at, w, h};
}

// MIPMap::Texel, mipmap.h:206-228 (a float texel is broadcast over the three lanes)
__device__ __forceinline__ V3f mipTexel(const MipLevel &L, int channels, int wrap, int s, int t) {
    if (wrap == IILE_WRAP_REPEAT) {
        s = texMod(s, L.w);
        t = texMod(t, L.h);
    } else if (wrap == IILE_WRAP_CLAMP) {
        s = min(max(s, 0), L.w - 1);
        t = min(max(t, 0), L.h - 1);
    } else if (s < 0 || s >= L.w || t < 0 || t >= L.h)
        return V3f(0, 0, 0);
    const float *p = L.texels + ((size_t)t * L.w + s) * channels;
    return channels == 3 ? V3f(p[0], p[1], p[2]) : V3f(p[0], p[0], p[0]);
}

// MIPMap::triangle, mipmap.h:258-270
static __device__ __noinline__ V3f mipTriangle(const DeviceScene &sc, const IileImage &im, int wrap, int level, float s0f,
                                        float t0f) {
    level = min(max(level, 0), im.nLevels - 1);
    const MipLevel L = mipLevel(sc, im, level);
    const float s = s0f * L.w - 0.5f, t = t0f * L.h - 0.5f;
    const int s0 = (int)floorf(s), t0 = (int)floorf(t);
    const float ds = s - s0, dt = t - t0;
    return (1 - ds) * (1 - dt) * mipTexel(L, im.channels, wrap, s0, t0) +
           (1 - ds) * dt * mipTexel(L, im.channels, wrap, s0, t0 + 1) +
           ds * (1 - dt) * mipTexel(L, im.channels, wrap, s0 + 1, t0) + ds * dt * mipTexel(L, im.channels, wrap, s0 + 1, t0 + 1);
}

__device__ __forceinline__ float texLog2(float x) {  // pbrt.h:325-328; std::log of a float: the host's logf
    return logLibm(x) * 1.442695040888963387004650940071f;
}

// MIPMap::EWA, mipmap.h:371-415.  sc.ewaWeights = MIPMap::weightLut (mipmap.h:197-203), filled by the host with
// the host's expf.
static __device__ __noinline__ V3f mipEWA(const DeviceScene &sc, const IileImage &im, int wrap, int level, float s, float t,
                                   float d0x, float d0y, float d1x, float d1y) {
    if (level >= im.nLevels) return mipTexel(mipLevel(sc, im, im.nLevels - 1), im.channels, wrap, 0, 0);
    const MipLevel L = mipLevel(sc, im, level);
    s = s * L.w - 0.5f;
    t = t * L.h - 0.5f;
    d0x *= L.w;
    d0y *= L.h;
    d1x *= L.w;
    d1y *= L.h;
    float A = d0y * d0y + d1y * d1y + 1;
    float B = -2 * (d0x * d0y + d1x * d1y);
    float C = d0x * d0x + d1x * d1x + 1;
    const float invF = 1 / (A * C - B * B * 0.25f);
    A *= invF;
    B *= invF;
    C *= invF;
    const float det = -B * B + 4 * A * C;
    const float invDet = 1 / det;
    const float uSqrt = sqrtf(det * C), vSqrt = sqrtf(A * det);
    const int s0 = (int)ceilf(s - 2 * invDet * uSqrt), s1 = (int)floorf(s + 2 * invDet * uSqrt);
    const int t0 = (int)ceilf(t - 2 * invDet * vSqrt), t1 = (int)floorf(t + 2 * invDet * vSqrt);
    V3f sum(0, 0, 0);
    float sumWts = 0;
    for (int it = t0; it <= t1; ++it) {
        const float tt = it - t;
        for (int is = s0; is <= s1; ++is) {
            const float ss = is - s;
            const float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
            if (r2 < 1) {
                const int index = min((int)(r2 * 128), 127);
                const float weight = sc.ewaWeights[index];
                sum = sum + mipTexel(L, im.channels, wrap, is, it) * weight;
                sumWts += weight;
            }
        }
    }
    return sdiv(sum, sumWts);
}

// MIPMap::Lookup(st, dst0, dst1), mipmap.h:239-256 and 343-369
static __device__ __noinline__ V3f mipLookup(const DeviceScene &sc, const IileTexture &tx, float s, float t, float d0x, float d0y,
                                      float d1x, float d1y) {
    const IileImage im = sc.images[tx.image];
    if (tx.trilinear) {
        const float width = 2 * fmaxf(fmaxf(fabsf(d0x), fabsf(d0y)), fmaxf(fabsf(d1x), fabsf(d1y)));
        const float level = im.nLevels - 1 + texLog2(fmaxf(width, 1e-8f));
        if (level < 0) return mipTriangle(sc, im, tx.wrap, 0, s, t);
        if (level >= im.nLevels - 1) return mipTexel(mipLevel(sc, im, im.nLevels - 1), im.channels, tx.wrap, 0, 0);
        const int iLevel = (int)floorf(level);
        const float delta = level - iLevel;
        return (1 - delta) * mipTriangle(sc, im, tx.wrap, iLevel, s, t) + delta * mipTriangle(sc, im, tx.wrap, iLevel + 1, s, t);
    }
    if (d0x * d0x + d0y * d0y < d1x * d1x + d1y * d1y) {
        float tmp = d0x;
        d0x = d1x, d1x = tmp;
        tmp = d0y;
        d0y = d1y, d1y = tmp;
    }
    const float majorLength = sqrtf(d0x * d0x + d0y * d0y);
    float minorLength = sqrtf(d1x * d1x + d1y * d1y);
    if (minorLength * tx.maxAnisotropy < majorLength && minorLength > 0) {
        const float scale = majorLength / (minorLength * tx.maxAnisotropy);
        d1x *= scale;
        d1y *= scale;
        minorLength *= scale;
    }
    if (minorLength == 0) return mipTriangle(sc, im, tx.wrap, 0, s, t);
    const float lod = fmaxf(0.f, im.nLevels - 1.f + texLog2(minorLength));
    const int ilod = (int)floorf(lod);
    const float w = lod - ilod;
    return (1 - w) * mipEWA(sc, im, tx.wrap, ilod, s, t, d0x, d0y, d1x, d1y) +
           w * mipEWA(sc, im, tx.wrap, ilod + 1, s, t, d0x, d0y, d1x, d1y);
}

// Texture::Evaluate for a non-scale texture at (u, v) with the interaction's uv differentials
static __device__ __noinline__ V3f evalLeafTexture(const DeviceScene &sc, const IileTexture &tx, float u, float v, const TexGeom &g) {
    if (tx.type == IILE_TEX_CONSTANT) return V3f(tx.value[0], tx.value[1], tx.value[2]);
    // UVMapping2D::Map, texture.cpp:93-99
    const float s = tx.su * u + tx.du, t = tx.sv * v + tx.dv;
    const V3f r = mipLookup(sc, tx, s, t, tx.su * g.dudx, tx.sv * g.dvdx, tx.su * g.dudy, tx.sv * g.dvdy);
    // convertOut (imagemap.h:107-112): RGB -> Spectrum is the identity for RGBSpectrum
    return r;
}
// Texture::Evaluate: ScaleTexture (scale.h:56-58) and MixTexture (mix.h:57-61) over their operands, at most DEPTH
// operators deep (operands sit earlier in the array: flat_io.cpp validates the DAG and its depth)
template <int DEPTH>
__device__ __forceinline__ V3f evalTextureD(const DeviceScene &sc, int index, float u, float v, const TexGeom &g) {
    const IileTexture tx = sc.textures[index];
    if constexpr (DEPTH > 0) {
        if (tx.type == IILE_TEX_SCALE)
            return evalTextureD<DEPTH - 1>(sc, tx.tex1, u, v, g) * evalTextureD<DEPTH - 1>(sc, tx.tex2, u, v, g);
        if (tx.type == IILE_TEX_MIX) {
            const V3f t1 = evalTextureD<DEPTH - 1>(sc, tx.tex1, u, v, g), t2 = evalTextureD<DEPTH - 1>(sc, tx.tex2, u, v, g);
            const float amt = tx.image >= 0 ? evalTextureD<DEPTH - 1>(sc, tx.image, u, v, g).x : tx.value[0];
            return (1 - amt) * t1 + amt * t2;
        }
        if (tx.type == IILE_TEX_CHECKERBOARD) {  // Checkerboard2DTexture::Evaluate, checkerboard.h:63-101
            const float s = tx.su * u + tx.du, t = tx.sv * v + tx.dv;  // UVMapping2D::Map, texture.cpp:93-99
            const bool first = ((int)floorf(s) + (int)floorf(t)) % 2 == 0;
            if (!tx.trilinear) {
                const float ds = fmaxf(fabsf(tx.su * g.dudx), fabsf(tx.su * g.dudy));
                const float dt = fmaxf(fabsf(tx.sv * g.dvdx), fabsf(tx.sv * g.dvdy));
                const float s0 = s - ds, s1 = s + ds, t0 = t - dt, t1 = t + dt;
                if (!(floorf(s0) == floorf(s1) && floorf(t0) == floorf(t1))) {
                    // the closed-form box filter over the checks the footprint straddles
                    auto bumpInt = [](float x) {
                        return (int)floorf(x / 2) + 2 * fmaxf(x / 2 - (int)floorf(x / 2) - 0.5f, 0.f);
                    };
                    const float sint = (bumpInt(s1) - bumpInt(s0)) / (2 * ds);
                    const float tint = (bumpInt(t1) - bumpInt(t0)) / (2 * dt);
                    float area2 = sint + tint - 2 * sint * tint;
                    if (ds > 1 || dt > 1) area2 = .5f;
                    return (1 - area2) * evalTextureD<DEPTH - 1>(sc, tx.tex1, u, v, g) +
                           area2 * evalTextureD<DEPTH - 1>(sc, tx.tex2, u, v, g);
                }
            }
            return evalTextureD<DEPTH - 1>(sc, first ? tx.tex1 : tx.tex2, u, v, g);
        }
    }
    return evalLeafTexture(sc, tx, u, v, g);
}
static __device__ __noinline__ V3f evalTexture(const DeviceScene &sc, int index, float u, float v, const TexGeom &g) {
    return evalTextureD<IILE_TEX_MAX_DEPTH>(sc, index, u, v, g);
}

// SurfaceInteraction::ComputeDifferentials, interaction.cpp:103-149 (n = the interaction's geometric normal)
// (dpdx / dpdy: the offsets to the auxiliary intersection points, what SpecularReflect / SpecularTransmit start
// the differentials of their rays from; optional)
__device__ __forceinline__ void computeDifferentials(V3f p, V3f n, V3f rxO, V3f rxD, V3f ryO, V3f ryD, TexGeom *g,
                                                     V3f *dpdx = nullptr, V3f *dpdy = nullptr) {
    g->dudx = g->dvdx = g->dudy = g->dvdy = 0;
    if (dpdx) *dpdx = *dpdy = V3f(0, 0, 0);
    const float d = dot(n, p);
    const float tx = -(dot(n, rxO) - d) / dot(n, rxD);
    if (isinf(tx) || isnan(tx)) return;
    const V3f px = rxO + tx * rxD;
    const float ty = -(dot(n, ryO) - d) / dot(n, ryD);
    if (isinf(ty) || isnan(ty)) return;
    const V3f py = ryO + ty * ryD;
    if (dpdx) *dpdx = px - p, *dpdy = py - p;
    int d0, d1;
    if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) d0 = 1, d1 = 2;
    else if (fabsf(n.y) > fabsf(n.z)) d0 = 0, d1 = 2;
    else d0 = 0, d1 = 1;
    auto comp = [](V3f a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); };
    const float A00 = comp(g->dpdu, d0), A01 = comp(g->dpdv, d0), A10 = comp(g->dpdu, d1), A11 = comp(g->dpdv, d1);
    const float Bx0 = comp(px, d0) - comp(p, d0), Bx1 = comp(px, d1) - comp(p, d1);
    const float By0 = comp(py, d0) - comp(p, d0), By1 = comp(py, d1) - comp(p, d1);
    const float det = A00 * A11 - A01 * A10;  // SolveLinearSystem2x2, transform.cpp:41-49
    if (fabsf(det) < 1e-10f) return;
    float x0 = (A11 * Bx0 - A01 * Bx1) / det, x1 = (A00 * Bx1 - A10 * Bx0) / det;
    if (!(isnan(x0) || isnan(x1))) g->dudx = x0, g->dvdx = x1;
    x0 = (A11 * By0 - A01 * By1) / det, x1 = (A00 * By1 - A10 * By0) / det;
    if (!(isnan(x0) || isnan(x1))) g->dudy = x0, g->dvdy = x1;
}

// the two offset rays of the camera ray through pFilm, in world space, scaled as Render scales them
// (hemiCam: the hemisphere's {cameraToWorld, worldToCamera} in a hemispheric pass, else null)
// pLens = lensRadius * ConcentricSampleDisk(sample.pLens) of the camera sample of (pixel, sample): drawn again from
// the two numbers generateKernel used (dimensions 3 and 4 of the stream), so that the offset rays start at the lens
// point's own bits (perspective.cpp:126-140 samples it a second time too)
__device__ __forceinline__ V3f lensPointOf(const RenderConsts &rc, uint32_t pixelIdx, uint32_t sIdx) {
    const uint32_t sbw = (uint32_t)(rc.sampleBounds[2] - rc.sampleBounds[0]);
    float lu, lv;
    if (rc.halton.tables) {
        const uint64_t index = haltonIndexForSample(rc.halton, rc.sampleBounds[0] + (int)(pixelIdx % sbw),
                                                    rc.sampleBounds[1] + (int)(pixelIdx / sbw), sIdx);
        lu = haltonScrambled(rc.halton.tables, index, 3);
        lv = haltonScrambled(rc.halton.tables, index, 4);
    } else if (rc.halton.sobol.matrices) {
        const uint64_t index = sobolIntervalToIndex(rc.halton.sobol, sIdx, pixelIdx % sbw, pixelIdx / sbw);
        lu = sobolSample(rc.halton.sobol.matrices, index, 3);
        lv = sobolSample(rc.halton.sobol.matrices, index, 4);
    } else {
        Pcg32 pcg;
        pcg.setSequence(((uint64_t)sIdx << 32) | pixelIdx);
        (void)pcg.uniformFloat();
        (void)pcg.uniformFloat();
        (void)pcg.uniformFloat();
        lu = pcg.uniformFloat();
        lv = pcg.uniformFloat();
    }
    float lx, ly;
    concentricSampleDisk(lu, lv, &lx, &ly);
    return V3f(lx * rc.lensRadius, ly * rc.lensRadius, 0);
}

// lensPt: the camera-space lens point of the main ray's sample (lensPointOf), or null where the caller no longer knows
// the sample — then it is recovered from the ray's origin through WorldToCamera, an ulp of a footprint off
__device__ __forceinline__ void cameraDifferentials(const RenderConsts &rc, const float *hemiCam, float fx, float fy, V3f o,
                                                    V3f d, V3f *rxO, V3f *rxD, V3f *ryO, V3f *ryD,
                                                    const V3f *lensPt = nullptr) {
    const float *m = hemiCam ? hemiCam : rc.cameraToWorld;
    auto vec = [&](V3f v) {
        return V3f(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
    };
    if (hemiCam) {
        // HemisphericCamera has no GenerateRayDifferential of its own: Camera's finite differences (camera.cpp:59-94)
        const float eps = (float).05;
        V3f co(0, 0, 0), cd = hemiCameraSpaceDir(rc.hemiSize, fx + eps, fy);
        rayToWorld(m, &co, &cd);
        *rxO = o + (co - o) / eps;
        *rxD = d + (cd - d) / eps;
        co = V3f(0, 0, 0), cd = hemiCameraSpaceDir(rc.hemiSize, fx, fy + eps);
        rayToWorld(m, &co, &cd);
        *ryO = o + (co - o) / eps;
        *ryD = d + (cd - d) / eps;
    } else if (rc.cameraType == IILE_CAMERA_PERSPECTIVE) {
        const float *r2c = rc.rasterToCamera;
        const V3f pCamera = xfPoint(r2c, V3f(fx, fy, 0));
        // dxCamera / dyCamera, perspective.cpp:58-62
        const V3f origin = xfPoint(r2c, V3f(0, 0, 0));
        const V3f dxCamera = xfPoint(r2c, V3f(1, 0, 0)) - origin, dyCamera = xfPoint(r2c, V3f(0, 1, 0)) - origin;
        if (rc.lensRadius > 0) {
            const V3f pLens = lensPt ? *lensPt : xfPoint(rc.worldToCamera, o);
            const V3f lens(pLens.x, pLens.y, 0);
            V3f dx = normalize(pCamera + dxCamera);
            float ft = rc.focalDistance / dx.z;
            *rxO = xfPoint(m, lens);
            *rxD = vec(normalize(ft * dx - lens));
            V3f dy = normalize(pCamera + dyCamera);
            ft = rc.focalDistance / dy.z;
            *ryO = xfPoint(m, lens);
            *ryD = vec(normalize(ft * dy - lens));
        } else {
            *rxO = *ryO = xfPoint(m, V3f(0, 0, 0));
            *rxD = vec(normalize(pCamera + dxCamera));
            *ryD = vec(normalize(pCamera + dyCamera));
        }
    } else if (rc.cameraType == IILE_CAMERA_ORTHOGRAPHIC) {  // orthographic.cpp:69-118 (dxCamera / dyCamera: :52-54)
        const float *r2c = rc.rasterToCamera;
        auto r2cVec = [&](V3f v) {
            return V3f(r2c[0] * v.x + r2c[1] * v.y + r2c[2] * v.z, r2c[4] * v.x + r2c[5] * v.y + r2c[6] * v.z,
                       r2c[8] * v.x + r2c[9] * v.y + r2c[10] * v.z);
        };
        const V3f dxCamera = r2cVec(V3f(1, 0, 0)), dyCamera = r2cVec(V3f(0, 1, 0));
        const V3f pCamera = xfPoint(r2c, V3f(fx, fy, 0));
        if (rc.lensRadius > 0) {
            const V3f pLens = lensPt ? *lensPt : xfPoint(rc.worldToCamera, o);
            const V3f lens(pLens.x, pLens.y, 0);
            const V3f dCam = normalize(pCamera + V3f(0, 0, rc.focalDistance) - lens);  // the main ray's camera-space d
            const float ft = rc.focalDistance / dCam.z;
            *rxO = *ryO = xfPoint(m, lens);
            *rxD = vec(normalize(pCamera + dxCamera + V3f(0, 0, ft) - lens));
            *ryD = vec(normalize(pCamera + dyCamera + V3f(0, 0, ft) - lens));
        } else {
            *rxO = xfPoint(m, pCamera + dxCamera);
            *ryO = xfPoint(m, pCamera + dyCamera);
            *rxD = *ryD = vec(V3f(0, 0, 1));
        }
    } else {
        // Camera::GenerateRayDifferential (camera.cpp:59-94): finite differences over a twentieth of a pixel
        const float eps = (float).05;
        V3f co, cd;
        cameraSpaceRay(rc, fx + eps, fy, 0.5f, 0.5f, &co, &cd);
        rayToWorld(m, &co, &cd);
        *rxO = o + (co - o) / eps;
        *rxD = d + (cd - d) / eps;
        cameraSpaceRay(rc, fx, fy + eps, 0.5f, 0.5f, &co, &cd);
        rayToWorld(m, &co, &cd);
        *ryO = o + (co - o) / eps;
        *ryD = d + (cd - d) / eps;
    }
    const float s = 1 / sqrtf((float)rc.spp);  // integrator.cpp:278-279
    *rxO = o + (*rxO - o) * s;
    *ryO = o + (*ryO - o) * s;
    *rxD = d + (*rxD - d) * s;
    *ryD = d + (*ryD - d) * s;
}

// the triangle's uv, dpdu / dpdv and shading geometry at the hit of the ray (o, d): the winning test repeated
static __device__ __forceinline__ void triangleTexGeom(const DeviceScene &sc, const IilePrim &pr, V3f o, V3f d, TexGeom *g) {
    RayCtx r;
    initRay(r, o.x, o.y, o.z, d.x, d.y, d.z, INFINITY);
    const float *P = sc.P;
    float t, b0, b1, b2;
    triangleTest(r, INFINITY, make_float3(P[3 * pr.v[0]], P[3 * pr.v[0] + 1], P[3 * pr.v[0] + 2]),
                 make_float3(P[3 * pr.v[1]], P[3 * pr.v[1] + 1], P[3 * pr.v[1] + 2]),
                 make_float3(P[3 * pr.v[2]], P[3 * pr.v[2] + 1], P[3 * pr.v[2] + 2]), &t, &b0, &b1, &b2);
    Surface again;
    triangleSurfaceInl(sc, pr, b0, b1, b2, d, &again, g);
    g->dudx = g->dudy = g->dvdx = g->dvdy = 0;
}

// The same for a sphere / disk / cylinder: Sphere::Intersect's parametric half with the second fundamental form
// (sphere.cpp:107-150, cylinder.cpp:108-150, disk.cpp:69-95) and the object->world SurfaceInteraction transform
// (transform.cpp:257-291); a quadric's shading geometry is its geometry.
static __device__ __noinline__ bool quadricTexGeom(const DeviceScene &sc, const IilePrim &pr, V3f o, V3f d, TexGeom *g) {
    const IileSphere &sp = sc.spheres[pr.v[0]];
    SphereHit h;
    float t;
    g->dudx = g->dudy = g->dvdx = g->dvdy = 0;
    if (!sphereQuadric(sp, o.x, o.y, o.z, d.x, d.y, d.z, INFINITY, &t, &h)) return false;
    const V3f pHit(h.px, h.py, h.pz);
    const float radius = sp.radius;
    const V3f dpdu(-sp.phiMax * pHit.y, sp.phiMax * pHit.x, 0);
    V3f dpdv, dndu(0, 0, 0), dndv(0, 0, 0);
    if (sp.flags & 4u) {  // disk: flat
        const float rHit = sqrtf(pHit.x * pHit.x + pHit.y * pHit.y);
        dpdv = (V3f(pHit.x, pHit.y, 0.f) * (radius - sp.zMax)) / rHit;
    } else {
        V3f d2Pduu = -sp.phiMax * sp.phiMax * V3f(pHit.x, pHit.y, 0), d2Pduv(0, 0, 0), d2Pdvv(0, 0, 0);
        if (sp.flags & 8u) {  // cylinder
            dpdv = V3f(0, 0, sp.zMax - sp.zMin);
        } else {
            const float theta = acosLibm(clampf(pHit.z / radius, -1, 1));
            const float zRadius = sqrtf(pHit.x * pHit.x + pHit.y * pHit.y);
            const float invZRadius = 1 / zRadius;
            const float cosPhi = pHit.x * invZRadius, sinPhi = pHit.y * invZRadius;
            const float dTheta = sp.thetaMax - sp.thetaMin;
            float sinTheta, cosThetaUnused;
            sincosLibm(theta, &sinTheta, &cosThetaUnused);
            dpdv = dTheta * V3f(pHit.z * cosPhi, pHit.z * sinPhi, -radius * sinTheta);
            d2Pduv = dTheta * pHit.z * sp.phiMax * V3f(-sinPhi, cosPhi, 0.f);
            d2Pdvv = -dTheta * dTheta * V3f(pHit.x, pHit.y, pHit.z);
        }
        const float E = dot(dpdu, dpdu), F = dot(dpdu, dpdv), G = dot(dpdv, dpdv);
        const V3f N = normalize(cross(dpdu, dpdv));
        const float e = dot(N, d2Pduu), f = dot(N, d2Pduv), gg = dot(N, d2Pdvv);
        const float invEGF2 = 1 / (E * G - F * F);
        dndu = (f * F - e * G) * invEGF2 * dpdu + (e * F - f * E) * invEGF2 * dpdv;
        dndv = (gg * F - f * G) * invEGF2 * dpdu + (f * F - gg * E) * invEGF2 * dpdv;
    }
    const float *m = sp.objectToWorld, *mi = sp.worldToObject;
    auto xfVec = [&](V3f v) {
        return V3f(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
                   m[8] * v.x + m[9] * v.y + m[10] * v.z);
    };
    auto xfNormal = [&](V3f v) {
        return V3f(mi[0] * v.x + mi[4] * v.y + mi[8] * v.z, mi[1] * v.x + mi[5] * v.y + mi[9] * v.z,
                   mi[2] * v.x + mi[6] * v.y + mi[10] * v.z);
    };
    g->u = h.u;
    g->v = h.v;
    g->dpdu = g->shDpdu = xfVec(dpdu);
    g->dpdv = g->shDpdv = xfVec(dpdv);
    g->shDndu = xfNormal(dndu);
    g->shDndv = xfNormal(dndv);
    return true;
}

// uv, dpdu / dpdv and the shading geometry of the hit of the ray (o, d) on primitive pr; *flipNormal: what
// SetShadingGeometry turns the recomputed shading normal by (reverseOrientation ^ transformSwapsHandedness)
static __device__ __forceinline__ void hitTexGeom(const DeviceScene &sc, const IilePrim &pr, V3f o, V3f d, TexGeom *g,
                                                  bool *flipNormal) {
    if (pr.type == IILE_PRIM_TRIANGLE) {
        triangleTexGeom(sc, pr, o, d, g);
        *flipNormal = (pr.flags & IILE_PRIM_FLIP_NORMAL) != 0;
    } else {
        quadricTexGeom(sc, pr, o, d, g);
        const uint32_t f = sc.spheres[pr.v[0]].flags;
        *flipNormal = ((f & 1u) != 0) ^ ((f & 2u) != 0);
    }
}

// Material::Bump and the material's texture slots, given the hit's geometry and uv differentials
static __device__ __noinline__ void applyTextures(const DeviceScene &sc, bool flipNormal, TexGeom &g, int bumpPasses,
                                                  Surface *s, IileMaterial *m) {
    const IileMaterialTextures mt = sc.materialTextures[m->texSet - 1];
    // Material::Bump first (every material's ComputeScatteringFunctions starts with it, e.g. matte.cpp:49)
    // bumpPasses: how often ComputeScatteringFunctions runs on this SurfaceInteraction — the IISPT runner calls it a
    // second time on the vertex find_intersection returned (iisptrenderrunner.cpp:689 and :536), and the second
    // Bump starts from the shading geometry the first one left (material.cpp:75-82 reads si->shading)
    if (mt.slot[IILE_TEXSLOT_BUMP] >= 0) {
        const int bt = mt.slot[IILE_TEXSLOT_BUMP];
        float du = .5f * (fabsf(g.dudx) + fabsf(g.dudy));
        if (du == 0) du = .0005f;
        const float uDisplace = evalTexture(sc, bt, g.u + du, g.v, g).x;
        float dv = .5f * (fabsf(g.dvdx) + fabsf(g.dvdy));
        if (dv == 0) dv = .0005f;
        const float vDisplace = evalTexture(sc, bt, g.u, g.v + dv, g).x;
        const float displace = evalTexture(sc, bt, g.u, g.v, g).x;
        for (int pass = 0; pass < bumpPasses; ++pass) {
            const V3f dpdu = g.shDpdu + (uDisplace - displace) / du * s->ns + displace * g.shDndu;
            const V3f dpdv = g.shDpdv + (vDisplace - displace) / dv * s->ns + displace * g.shDndv;
            // SetShadingGeometry(dpdu, dpdv, ..., false), interaction.cpp:73-92
            V3f shn = normalize(cross(dpdu, dpdv));
            if (flipNormal) shn = -shn;
            shn = faceforward(shn, s->n);
            s->ns = shn;
            s->ss = normalize(dpdu);  // BSDF ctor, reflection.h:153-161
            s->ts = cross(s->ns, s->ss);
            g.shDpdu = dpdu;
            g.shDpdv = dpdv;
        }
    }
    auto spectrum = [&](int slot, float *dst) {
        if (mt.slot[slot] < 0) return;
        const V3f c = evalTexture(sc, mt.slot[slot], g.u, g.v, g);
        dst[0] = c.x, dst[1] = c.y, dst[2] = c.z;
    };
    auto scalar = [&](int slot, float *dst) {
        if (mt.slot[slot] >= 0) *dst = evalTexture(sc, mt.slot[slot], g.u, g.v, g).x;
    };
    spectrum(IILE_TEXSLOT_KD, m->Kd);
    spectrum(IILE_TEXSLOT_KS, m->Ks);
    spectrum(IILE_TEXSLOT_KT, m->Kt);
    spectrum(IILE_TEXSLOT_KR, m->Kr);
    spectrum(IILE_TEXSLOT_OPACITY, m->opacity);
    scalar(IILE_TEXSLOT_SIGMA, &m->sigma);
    scalar(IILE_TEXSLOT_ROUGHNESS, &m->roughness);
    scalar(IILE_TEXSLOT_UROUGHNESS, &m->uRoughness);
    scalar(IILE_TEXSLOT_VROUGHNESS, &m->vRoughness);
}

// The surface of a hit on a triangle whose material has textures: evaluates the material's texture slots into
// *m (a copy of the constant record) and applies the bump map to the shading frame of *s.  (o, d) is the ray
// that found the hit; cameraRay says whether it carries differentials (bounce 0 of a path).
static __device__ __noinline__ void resolveTextured(const DeviceScene &sc, const RenderConsts &rc, int primIdx, V3f o, V3f d,
                                                bool cameraRay, float pFilmX, float pFilmY, const float *hemiCam,
                                                int bumpPasses, Surface *s, IileMaterial *m, const V3f *lensPt = nullptr) {
    const IilePrim pr = sc.prims[primIdx];
    TexGeom g;
    bool flipNormal;
    hitTexGeom(sc, pr, o, d, &g, &flipNormal);
    if (cameraRay) {
        V3f rxO, rxD, ryO, ryD;
        cameraDifferentials(rc, hemiCam, pFilmX, pFilmY, o, d, &rxO, &rxD, &ryO, &ryD, lensPt);
        computeDifferentials(s->p, s->n, rxO, rxD, ryO, ryD, &g);
    }
    applyTextures(sc, flipNormal, g, bumpPasses, s, m);
}

// The direct integrators' vertex (DirectLightingIntegrator / WhittedIntegrator::Li): rays carry differentials through
// the specular recursion (SpecularReflect / SpecularTransmit, integrator.cpp:341-432), so the differentials are
// given, every triangle vertex needs them (a mirror hands them on), and what the two functions read from the
// SurfaceInteraction comes back: dpdx, dpdy, dndx = shading.dndu * dudx + shading.dndv * dvdx, dndy.
static __device__ __noinline__ void resolveDirectVertex(const DeviceScene &sc, int primIdx, V3f o, V3f d, bool hasDiffs,
                                                        V3f rxO, V3f rxD, V3f ryO, V3f ryD, Surface *s, IileMaterial *m,
                                                        V3f *dpdx, V3f *dpdy, V3f *dndx, V3f *dndy) {
    const IilePrim pr = sc.prims[primIdx];
    TexGeom g;
    bool flipNormal;
    hitTexGeom(sc, pr, o, d, &g, &flipNormal);
    *dpdx = *dpdy = V3f(0, 0, 0);
    if (hasDiffs) computeDifferentials(s->p, s->n, rxO, rxD, ryO, ryD, &g, dpdx, dpdy);
    *dndx = g.shDndu * g.dudx + g.shDndv * g.dvdx;
    *dndy = g.shDndu * g.dudy + g.shDndv * g.dvdy;
    if (m->texSet) applyTextures(sc, flipNormal, g, 1, s, m);
}

}  // namespace iile
