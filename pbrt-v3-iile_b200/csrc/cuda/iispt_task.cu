// One task of IisptRenderRunner::run (src/integrators/iisptrenderrunner.cpp:216-596) on the device, the two
// stages either side of the hemisphere batch (RenderView + the network, render_api.cu / hemi_maps.cu):
//
//   hemi points  the task's grid of "hemi points" (run, :252-389): camera sample, GenerateRayDifferential,
//                find_intersection (:648-751: follow specular bounces, stop at the first non-specular
//                vertex) -> position and axis of the auxiliary HemisphericCamera (normal turned against the
//                ray, SpawnRay origin)
//   gather       every film pixel of the task (run, :404-570): camera sample, find_intersection,
//                compute_fpixel_weights (:939-1019, tools/iisptmathutils.h), the BSDF, and sample_hemisphere /
//                estimate_direct (:16-178) against the network's films of the four neighbouring hemi points
//
// find_intersection is a small wavefront of its own: the camera rays are cast by the closest-hit kernel of
// the ray-cast API (cast_kernels.cu), vertexKernel consumes the hits, and the few paths that met a specular
// surface go round again (at most 24 times, :663).
//
// Random numbers.  The reference draws the camera samples and the BSDF samples of find_intersection from
// its Sampler after sampler_next_pixel() (:917-933: every camera sample gets a fresh "pixel" (counter, 0)),
// and sample_hemisphere's numbers from ONE serial PCG32 per thread (IisptRng, tools/iisptrng.cpp).  A serial
// stream cannot be shared by the pixels of a kernel; as for the path integrator (DESIGN.md "Sample
// streams") the device uses one PCG32 stream per camera sample — sequence = samplerSample << 32 | counter,
// what the harness's PixelStreamSampler draws from — and one per pixel for the gather — IisptRng(rngSeed +
// pixel index inside the task), which oracle/ref_harness/iispt_task.cpp installs before each pixel's call of
// the reference's own sample_hemisphere.
#include "camera.cuh"
#include "cast_kernels.h"
#include "intersect.cuh"
#include "pcg.cuh"
#include "sphere.cuh"
#include "texture.cuh"
#include "wavefront.h"

namespace iile {

namespace {

constexpr int kBlock = 128;
constexpr int kImportanceSamples = 16;  // IisptRenderRunner::HEMISPHERIC_IMPORTANCE_SAMPLES, iisptrenderrunner.h:33

__device__ __forceinline__ V3f f4(float4 v) { return V3f(v.x, v.y, v.z); }

// hemi grid of a task (run, :252-389): x0, x0 + tilesize, ... and always x1 - 1
__host__ __device__ inline int gridCount(int lo, int hi, int tilesize) {
    const int span = hi - 1 - lo;
    return span <= 0 ? 1 : (span + tilesize - 1) / tilesize + 1;
}
__device__ __forceinline__ int gridCoord(int lo, int hi, int tilesize, int k) { return min(lo + k * tilesize, hi - 1); }
__device__ __forceinline__ int gridIndex(int lo, int hi, int tilesize, int c) {
    return c == hi - 1 ? gridCount(lo, hi, tilesize) - 1 : (c - lo) / tilesize;
}

// SurfaceInteraction of a hit given by its scene-order primitive (IileHit::prim): the winning test is
// repeated on the same inputs (same bits), as wavefront.cu's surfaceAt does from the leaf slot.
__device__ __noinline__ bool surfaceAtPrim(const DeviceScene &sc, int primIdx, V3f o, V3f d, Surface *s) {
    const IilePrim pr = sc.prims[primIdx];
    if (pr.type == IILE_PRIM_TRIANGLE) {
        RayCtx r;
        initRay(r, o.x, o.y, o.z, d.x, d.y, d.z, INFINITY);
        const float *P = sc.P;
        float t, b0, b1, b2;
        if (!triangleTest(r, INFINITY, make_float3(P[3 * pr.v[0]], P[3 * pr.v[0] + 1], P[3 * pr.v[0] + 2]),
                          make_float3(P[3 * pr.v[1]], P[3 * pr.v[1] + 1], P[3 * pr.v[1] + 2]),
                          make_float3(P[3 * pr.v[2]], P[3 * pr.v[2] + 1], P[3 * pr.v[2] + 2]), &t, &b0, &b1, &b2))
            return false;
        triangleSurface(sc, pr, b0, b1, b2, d, s);
    } else {
        const IileSphere &sp = sc.spheres[pr.v[0]];
        SphereHit sh;
        float t;
        if (!sphereQuadric(sp, o.x, o.y, o.z, d.x, d.y, d.z, INFINITY, &t, &sh)) return false;
        sphereSurface(sp, pr, sh, s);
    }
    return true;
}

// work item -> pixel of the main camera and the sampler's pixel counter
__device__ __forceinline__ void itemPixel(const IisptTaskDev &t, uint32_t w, int *px, int *py, uint32_t *counter) {
    if (t.stage == 0) {  // hemi points, rows of the grid (run :252-389 advances x first)
        *px = gridCoord(t.x0, t.x1, t.tilesize, (int)(w % (uint32_t)t.gridW));
        *py = gridCoord(t.y0, t.y1, t.tilesize, (int)(w / (uint32_t)t.gridW));
        *counter = t.samplerCounter + 1u + w;
    } else {  // film pixels, row-major (run :416-417); the hemi points took the first gridW * gridH counters
        const uint32_t width = (uint32_t)(t.x1 - t.x0);
        *px = t.x0 + (int)(w % width);
        *py = t.y0 + (int)(w / width);
        *counter = t.samplerCounter + 1u + (uint32_t)(t.gridW * t.gridH) + w;
    }
}

// sampler_next_pixel + GetCameraSample + GenerateRayDifferential (run :262-275 / :474-484)
__global__ void __launch_bounds__(kBlock) startKernel(RenderConsts rc, IisptTaskDev t, uint32_t n, IisptVertex *verts,
                                                      IileRay *rays, uint32_t *active) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    int px, py;
    uint32_t counter;
    itemPixel(t, w, &px, &py, &counter);
    Pcg32 g;
    g.setSequence(((uint64_t)t.samplerSample << 32) | counter);
    const float fx = (float)px + g.uniformFloat();
    const float fy = (float)py + g.uniformFloat();
    (void)g.uniformFloat();  // time
    const float lu = g.uniformFloat(), lv = g.uniformFloat();
    V3f o, d;
    cameraSpaceRay(rc, fx, fy, lu, lv, &o, &d);
    rayToWorld(rc.cameraToWorld, &o, &d);
    IisptVertex v;
    v.o = make_float4(o.x, o.y, o.z, fx);  // pFilm rides along for the camera ray's differentials (texture stage)
    v.d = make_float4(d.x, d.y, d.z, fy);
    v.beta = make_float4(1, 1, 1, 0);
    v.aux = make_float4(0, 0, 0, 0);
    v.rngState = g.state;
    v.rngInc = g.inc;
    v.status = kIisptActive;
    v.prim = -1;
    v.pad[0] = v.pad[1] = 0;
    verts[w] = v;
    rays[w] = IileRay{{o.x, o.y, o.z}, INFINITY, {d.x, d.y, d.z}, 0.f};
    active[w] = w;
}

// The texture stage of the task (scenes with image textures or bump maps only; cf. wavefront.cu's textureKernel):
// the material of a hit with its textures evaluated, and the bump-mapped shading normal and tangent, left in
// texRecord[8 * item] for the kernel that builds the BSDF.  Camera rays carry GenerateRayDifferential's offsets
// with ScaleDifferentials(1.0) (run :266-271, :494-499: rc.spp is 1 here); the rays find_intersection spawns carry
// none (:742).  gather = 0: before vertexKernel, over the active items of this bounce; gather = 1: over the
// vertices find_intersection returned, for the SECOND ComputeScatteringFunctions on them (:536) — its Bump
// starts from the shading geometry the first one left.
__global__ void __launch_bounds__(kBlock) iisptTextureKernel(DeviceScene sc, RenderConsts rc, IisptTaskDev t, uint32_t nItems,
                                                             const uint32_t *activeIn, const IileHit *hits,
                                                             const IisptVertex *verts, float4 *texRecord, int bounce,
                                                             int gather) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nItems) return;
    const uint32_t w = gather ? k : activeIn[k];
    const IisptVertex v = verts[w];
    int prim;
    bool cameraRay;
    if (gather) {
        if (v.status != kIisptFound) return;
        prim = v.prim;
        cameraRay = v.pad[0] == 0;
    } else {
        prim = hits[k].prim;
        cameraRay = bounce == 0;
    }
    if (prim < 0) return;
    const IilePrim pr = sc.prims[prim];
    if (pr.material < 0 || !sc.materials[pr.material].texSet) return;
    Surface s;
    if (!surfaceAtPrim(sc, prim, f4(v.o), f4(v.d), &s)) return;
    IileMaterial m = sc.materials[pr.material];
    // the lens point of the item's camera sample, drawn again as startKernel drew it (texture.cuh::cameraDifferentials)
    V3f lensPt(0, 0, 0);
    const bool lens = cameraRay && rc.lensRadius > 0;
    if (lens) {
        int px, py;
        uint32_t counter;
        itemPixel(t, w, &px, &py, &counter);
        Pcg32 g;
        g.setSequence(((uint64_t)t.samplerSample << 32) | counter);
        (void)g.uniformFloat();
        (void)g.uniformFloat();
        (void)g.uniformFloat();
        const float lu = g.uniformFloat(), lv = g.uniformFloat();
        float lx, ly;
        concentricSampleDisk(lu, lv, &lx, &ly);
        lensPt = V3f(lx * rc.lensRadius, ly * rc.lensRadius, 0);
    }
    resolveTextured(sc, rc, prim, f4(v.o), f4(v.d), cameraRay, v.o.w, v.d.w, nullptr, gather ? 2 : 1, &s, &m,
                    lens ? &lensPt : nullptr);
    float4 *rec = texRecord + 8 * (size_t)w;
    const float4 *src = reinterpret_cast<const float4 *>(&m);
#pragma unroll
    for (int j = 0; j < 6; ++j) rec[j] = src[j];
    rec[6] = make_float4(s.ns.x, s.ns.y, s.ns.z, 0);
    rec[7] = make_float4(s.ss.x, s.ss.y, s.ss.z, 0);
}

// the material the BSDF is built from: the texture stage's record when the material has textures
__device__ __forceinline__ const IileMaterial *materialFor(const DeviceScene &sc, const float4 *texRecord, uint32_t w,
                                                           Surface *s) {
    const IileMaterial *mp = &sc.materials[s->material];
    if (mp->texSet && texRecord) {
        const float4 *rec = texRecord + 8 * (size_t)w;
        mp = reinterpret_cast<const IileMaterial *>(rec);
        s->ns = f4(rec[6]);
        s->ss = f4(rec[7]);
        s->ts = cross(s->ns, s->ss);
    }
    return mp;
}

// One iteration of find_intersection's loop (:663-745) for the paths still looking for their vertex.
__global__ void __launch_bounds__(kBlock) vertexKernel(DeviceScene sc, uint32_t nActive, const uint32_t *activeIn,
                                                       const IileHit *hits, IisptVertex *verts, uint32_t *activeOut,
                                                       IileRay *raysOut, uint32_t *counter, int bounce,
                                                       const float4 *texRecord) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nActive) return;
    const uint32_t w = activeIn[k];
    IisptVertex v = verts[w];
    const IileHit hit = hits[k];
    const V3f o = f4(v.o), d = f4(v.d);
    Surface s;
    if (hit.prim < 0 || !surfaceAtPrim(sc, hit.prim, o, d, &s)) {
        verts[w].status = kIisptNoIntersection;  // :672-679 (the background is the direct pass's business)
        return;
    }
    V3f next;
    V3f beta = f4(v.beta);
    if (s.material < 0) {
        next = d;  // no BSDF: skip the interface, :690-694
    } else {
        Bsdf bsdf;
        makeBsdf(*materialFor(sc, texRecord, w, &s), s, &bsdf, true);  // ComputeScatteringFunctions(ray, arena, true), :689
        Pcg32 g;
        g.state = v.rngState;
        g.inc = v.rngInc;
        const float u0 = g.uniformFloat(), u1 = g.uniformFloat();  // sampler->Get2D()
        verts[w].rngState = g.state;
        V3f wi;
        float pdf;
        int sampled;
        const V3f f = bsdfSample(bsdf, -d, &wi, u0, u1, &pdf, BSDF_ALL, &sampled);  // wo = -ray.d, :707 (not isect.wo)
        if (isBlack(f) || pdf == 0.f) {  // :712-715
            verts[w].status = kIisptBlack;
            return;
        }
        if (!(sampled & BSDF_SPECULAR)) {  // :718-726: IISPT takes over from this vertex
            const V3f nTurned = dot(s.n, d) > 0.f ? -s.n : s.n;  // run :301-311, compute_fpixel_weights :952-962
            const V3f auxO = offsetRayOrigin(s.p, s.pError, s.n, nTurned);  // isect.SpawnRay(Vector3f(surface_normal))
            verts[w].status = kIisptFound;
            verts[w].prim = hit.prim;
            verts[w].pad[0] = bounce;  // 0: the ray that found the vertex is the camera ray (differentials)
            verts[w].aux = make_float4(auxO.x, auxO.y, auxO.z, 0);
            verts[w].n = make_float4(nTurned.x, nTurned.y, nTurned.z, 0);
            verts[w].p = make_float4(s.p.x, s.p.y, s.p.z, 0);
            return;
        }
        beta = beta * sdiv(f * absDot(wi, s.ns), pdf);  // :729
        const float by = luminance(beta);
        if (by < 0.f || isnan(by)) {  // :731-734
            verts[w].status = kIisptBlack;
            return;
        }
        next = wi;
    }
    if (bounce == 23) {  // :748-750: depth 24 reached
        verts[w].status = kIisptBlack;
        return;
    }
    const V3f no = offsetRayOrigin(s.p, s.pError, s.n, next);
    verts[w].o = make_float4(no.x, no.y, no.z, 0);
    verts[w].d = make_float4(next.x, next.y, next.z, 0);
    verts[w].beta = make_float4(beta.x, beta.y, beta.z, 0);
    const uint32_t at = atomicAdd(counter, 1u);
    activeOut[at] = w;
    raysOut[at] = IileRay{{no.x, no.y, no.z}, INFINITY, {next.x, next.y, next.z}, 0.f};
}

// run :293-325: a hemi point exists where an intersection was found and beta.y() > 0
__global__ void __launch_bounds__(kBlock) hemiPointsKernel(uint32_t n, const IisptVertex *verts, float *pos, float *dir,
                                                           uint8_t *valid) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const IisptVertex v = verts[w];
    const bool ok = v.status == kIisptFound && luminance(f4(v.beta)) > 0.f;
    valid[w] = ok ? 1 : 0;
    for (int k = 0; k < 3; ++k) {
        pos[3 * w + k] = ok ? (&v.aux.x)[k] : 0.f;
        dir[3 * w + k] = ok ? (&v.n.x)[k] : 0.f;
    }
}

// RNG::UniformUInt32(b), rng.h:95-105
__device__ __forceinline__ uint32_t uniformBelow(Pcg32 &g, uint32_t b) {
    const uint32_t threshold = (~b + 1u) % b;
    for (;;) {
        const uint32_t r = g.uniformUInt32();
        if (r >= threshold) return r % b;
    }
}

// IntensityFilm::get_camera_coord_jacobian (film/intensityfilm.cpp:60-66) of hemi camera `film`
__device__ __forceinline__ V3f nnPixelJacobian(const float *films, int film, int size, int x, int y) {
    const float *px = films + (((size_t)film * size + (size - 1 - y)) * size + x) * 3;
    const float absVertical = (float)y / (float)size;
    const float polar = (float)(3.14159265358979323846 * (double)absVertical);
    const float jacobian = sinGlibc(polar);  // `sin(polar)` of a float: the host's sinf (libm_exact.cuh)
    return V3f(px[0] * jacobian, px[1] * jacobian, px[2] * jacobian);
}

// estimate_direct, iisptrenderrunner.cpp:16-143.  cam: {CameraToWorld.m, WorldToCamera.m} of the hemi camera.
__device__ __noinline__ V3f estimateDirect(const Bsdf &bsdf, const Surface &s, int rx, int ry, const float *cam,
                                           const float *films, int film, int size, Pcg32 &g) {
    constexpr int flags = BSDF_ALL & ~BSDF_SPECULAR;
    const float lightPdf = (float)(1.0 / 6.28);
    const float bsdfRatio = 0.4394f, emRatio = 1.098f;
    V3f Ld(0, 0, 0);
    // HemisphericCamera::get_light_sample_nn, cameras/hemispheric.cpp:89-105
    V3f wi;
    {
        const float theta = IILE_PI * (float)ry / (float)size;
        const float phi = IILE_PI * (float)rx / (float)size;
        float sT, cT, sP, cP;
        sincosLibm(theta, &sT, &cT);
        sincosLibm(phi, &sP, &cP);
        const V3f dc(sT * cP, cT, sT * sP);
        wi = V3f(cam[0] * dc.x + cam[1] * dc.y + cam[2] * dc.z, cam[4] * dc.x + cam[5] * dc.y + cam[6] * dc.z,
                 cam[8] * dc.x + cam[9] * dc.y + cam[10] * dc.z);
    }
    const V3f Li = nnPixelJacobian(films, film, size, rx, ry);
    if (!isBlack(Li)) {
        const V3f f = bsdfF(bsdf, s.wo, wi, flags) * absDot(wi, s.ns);
        const float scatteringPdf = bsdfPdf(bsdf, s.wo, wi, flags);
        if (!isBlack(f)) {
            const float weight = powerHeuristic(lightPdf, scatteringPdf);
            Ld = Ld + sdiv(((emRatio * f) * Li) * weight, lightPdf);
        }
    }
    {
        // Point2f uScattering(rng->uniform_float(), rng->uniform_float()): g++ evaluates the arguments of
        // the constructor call right to left, so the FIRST number drawn is y (pinned by the parity test)
        const float u1 = g.uniformFloat(), u0 = g.uniformFloat();
        float scatteringPdf;
        int sampled;
        V3f f = bsdfSample(bsdf, s.wo, &wi, u0, u1, &scatteringPdf, flags, &sampled);
        if (scatteringPdf > 0) f = f * absDot(wi, s.ns);
        if (!isBlack(f) && scatteringPdf > 0) {
            const float weight = (sampled & BSDF_SPECULAR) ? 1.f : powerHeuristic(scatteringPdf, lightPdf);
            // HemisphericCamera::getLightSampleNn, cameras/hemispheric.cpp:44-60
            const float *m = cam + 16;
            const V3f wc(m[0] * wi.x + m[1] * wi.y + m[2] * wi.z, m[4] * wi.x + m[5] * wi.y + m[6] * wi.z,
                         m[8] * wi.x + m[9] * wi.y + m[10] * wi.z);
            const float theta = acosGlibc(wc.y);
            const float phi = atan2Glibc(wc.z, wc.x);
            const float fyv = (float)size * theta / IILE_PI, fxv = (float)size * phi / IILE_PI;
            // (int) of a NaN is INT_MIN on the reference's x86 (fails the range test); truncation otherwise
            if (!isnan(fyv) && !isnan(fxv)) {
                const int y = (int)fyv, x = (int)fxv;
                if (x >= 0 && x < size && y >= 0 && y < size) {
                    const V3f Lb = nnPixelJacobian(films, film, size, x, y);
                    if (!isBlack(Lb)) Ld = Ld + sdiv(((bsdfRatio * f) * Lb) * weight, scatteringPdf);
                }
            }
        }
    }
    return Ld;
}

// The pixel loop of run (:416-570) for the pixels whose find_intersection ended at a vertex.
// hemiPos / hemiDir / hemiFilm: per hemi point of the task's grid (hemiFilm < 0: no camera there);
// cams: per film {CameraToWorld.m, WorldToCamera.m}; films: [nFilms][size][size][3] in ImageFilm order.
__global__ void __launch_bounds__(kBlock) gatherKernel(DeviceScene sc, RenderConsts rc, IisptTaskDev t, uint32_t n,
                                                       const IisptVertex *verts, const float *hemiPos, const float *hemiDir,
                                                       const int *hemiFilm, const float *cams, const float *films,
                                                       float4 *out, const float4 *texRecord) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const IisptVertex v = verts[w];
    float4 result = make_float4(0, 0, 0, 0);
    const V3f beta = f4(v.beta);
    Surface s;
    if (v.status == kIisptFound && luminance(beta) > 0.f && surfaceAtPrim(sc, v.prim, f4(v.o), f4(v.d), &s) &&
        s.material >= 0) {
        const int width = t.x1 - t.x0;
        const int fx = t.x0 + (int)(w % (uint32_t)width), fy = t.y0 + (int)(w / (uint32_t)width);
        // the four neighbouring hemi points, in run's order S, E, R, B (:423-455)
        const int sx = fx - (fx - t.x0) % t.tilesize, sy = fy - (fy - t.y0) % t.tilesize;
        const int ex = min(sx + t.tilesize, t.x1 - 1), ey = min(sy + t.tilesize, t.y1 - 1);
        const int nx[4] = {sx, ex, ex, sx}, ny[4] = {sy, ey, sy, ey};
        // Camera::getCameraWorldPosition (core/camera.cpp:115-124): the origin of the ray through pFilm (0, 0)
        V3f camO, camD;
        cameraSpaceRay(rc, 0.f, 0.f, 0.f, 0.f, &camO, &camD);
        rayToWorld(rc.cameraToWorld, &camO, &camD);
        // compute_fpixel_weights, :939-1019
        const V3f auxD = f4(v.n), isectP = f4(v.p);
        int film[4];
        float weight[4];
        for (int i = 0; i < 4; ++i) {
            const int h = gridIndex(t.y0, t.y1, t.tilesize, ny[i]) * t.gridW + gridIndex(t.x0, t.x1, t.tilesize, nx[i]);
            film[i] = hemiFilm[h];
            // iispt::weighting_distance_positions
            float dx2 = (float)(fx - nx[i]);
            dx2 = dx2 * dx2;
            float dy2 = (float)(fy - ny[i]);
            dy2 = dy2 * dy2;
            const float pdist = sqrtf(dx2 + dy2);
            float wdpos = (float)t.tilesize != 0.f ? pdist / (float)t.tilesize : pdist;
            wdpos = wdpos < 0.f ? 0.f : (wdpos > 1.f ? 1.f : wdpos);
            float wdnor = 0.f, wdd = 0.f;
            if (film[i] >= 0) {
                // iispt::weighting_distance_normals
                V3f a = auxD, b = ld3(hemiDir + 3 * h);
                const float al = length(a), bl = length(b);
                if (al <= 0.f || bl <= 0.f) wdnor = 1.f;
                else {
                    a = a / al;
                    b = b / bl;
                    const float dt = dot(a, b);
                    wdnor = dt < 0.f ? 1.f : 1.f - dt;
                }
                // iispt::weightingCameraDistance
                const float toCamera = length(camO - isectP);
                if ((double)toCamera < 1e-10) wdd = 0.f;
                else {
                    const float sampleToCamera = length(camO - ld3(hemiPos + 3 * h));
                    const float relativeError = fabsf(toCamera - sampleToCamera) / toCamera * 1.f;
                    const float wt = 1.0f - relativeError;
                    wdd = wt < 0.f ? 0.f : (wt > 1.f ? 1.f : wt);
                }
            }
            const float wod = wdpos * wdnor + wdpos * wdd + wdpos;
            weight[i] = (float)(fmax(0.0, 2.0 - (double)wod) + 0.001);
        }
        {  // iispt::weights_to_probabilities
            float total = 0.f;
            for (int i = 0; i < 4; ++i) total += weight[i];
            if (total > 0.f)
                for (int i = 0; i < 4; ++i) weight[i] = weight[i] / total;
        }
        Bsdf bsdf;
        makeBsdf(*materialFor(sc, texRecord, w, &s), s, &bsdf, false);  // f_isect.ComputeScatteringFunctions(f_ray, arena), :536
        // sample_hemisphere, :147-178
        Pcg32 g;
        g.setSequence(t.rngSeed + w);
        V3f L(0, 0, 0);
        int taken = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < kImportanceSamples; ++j) {
                const float rr = g.uniformFloat();
                if (rr < weight[i]) {
                    ++taken;
                    if (film[i] >= 0) {
                        const int rx = (int)uniformBelow(g, (uint32_t)t.hemiSize);
                        const int ry = (int)uniformBelow(g, (uint32_t)t.hemiSize);
                        L = L + estimateDirect(bsdf, s, rx, ry, cams + 32 * (size_t)film[i], films, film[i], t.hemiSize, g);
                    }
                }
            }
        if (taken > 0) {
            L = sdiv(L, (float)taken);
            const V3f c = beta * L;  // additions_spectrum = f_beta * L, weight 0.5 (:563-565)
            result = make_float4(c.x, c.y, c.z, 0.5f);
        } else
            result = make_float4(0, 0, 0, 0.5f);
    }
    out[w] = result;
}

}  // namespace

void iisptGrid(const IisptTaskDev &t, int *gridW, int *gridH) {
    *gridW = gridCount(t.x0, t.x1, t.tilesize);
    *gridH = gridCount(t.y0, t.y1, t.tilesize);
}

// find_intersection for the n work items of the task's stage; leaves verts[n] filled.  Synchronises the
// stream once per bounce (the number of paths still on a specular chain is read back).
// r.ScaleDifferentials(1.0), run :271 / :499: the texture stage scales the offsets by 1 / sqrt(rc.spp)
static RenderConsts taskConsts(const RenderConsts &rc) {
    RenderConsts one = rc;
    one.spp = 1;
    return one;
}

cudaError_t iisptFindVertices(const DeviceScene &sc, const RenderConsts &rc, const IisptTaskDev &t, uint32_t n,
                              IisptScratch scratch, cudaStream_t stream, uint64_t *launches) {
    const uint32_t blocks = (n + kBlock - 1) / kBlock;
    startKernel<<<blocks, kBlock, 0, stream>>>(rc, t, n, scratch.verts, scratch.rays[0], scratch.active[0]);
    ++*launches;
    uint32_t nActive = n;
    for (int bounce = 0; bounce < 24 && nActive > 0; ++bounce) {  // :663
        const int cur = bounce & 1;
        cudaError_t e = cudaMemsetAsync(scratch.counter, 0, sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
        e = launchClosestHit(sc, scratch.rays[cur], nActive, scratch.hits, stream);
        if (e != cudaSuccess) return e;
        if (scratch.texRecord) {
            iisptTextureKernel<<<(nActive + kBlock - 1) / kBlock, kBlock, 0, stream>>>(
                sc, taskConsts(rc), t, nActive, scratch.active[cur], scratch.hits, scratch.verts, scratch.texRecord, bounce, 0);
            ++*launches;
        }
        vertexKernel<<<(nActive + kBlock - 1) / kBlock, kBlock, 0, stream>>>(sc, nActive, scratch.active[cur], scratch.hits,
                                                                            scratch.verts, scratch.active[cur ^ 1],
                                                                            scratch.rays[cur ^ 1], scratch.counter, bounce,
                                                                            scratch.texRecord);
        *launches += 2;
        e = cudaMemcpyAsync(&nActive, scratch.counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

cudaError_t iisptHemiPoints(uint32_t n, const IisptVertex *verts, float *pos, float *dir, uint8_t *valid,
                            cudaStream_t stream) {
    hemiPointsKernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(n, verts, pos, dir, valid);
    return cudaGetLastError();
}

cudaError_t iisptGather(const DeviceScene &sc, const RenderConsts &rc, const IisptTaskDev &t, uint32_t n,
                        const IisptVertex *verts, const float *hemiPos, const float *hemiDir, const int *hemiFilm,
                        const float *cams, const float *films, float4 *out, float4 *texRecord, cudaStream_t stream) {
    if (texRecord)
        iisptTextureKernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(sc, taskConsts(rc), t, n, nullptr, nullptr, verts,
                                                                            texRecord, 0, 1);
    gatherKernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(sc, rc, t, n, verts, hemiPos, hemiDir, hemiFilm, cams,
                                                                  films, out, texRecord);
    return cudaGetLastError();
}

}  // namespace iile
