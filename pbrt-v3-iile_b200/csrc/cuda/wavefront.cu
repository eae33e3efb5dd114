// Wavefront PathIntegrator: the SamplerIntegrator::Render tile loop
// (src/core/integrator.cpp:227-339) and PathIntegrator::Li's bounce loop
// (src/integrators/path.cpp:64-194) re-expressed as a sequence of kernels over
// queues of path indices, one path per (pixel, sample):
//
//   generate   RandomSampler::GetCameraSample + PerspectiveCamera::GenerateRayDifferential
//              (sampler.cpp:46-52, perspective.cpp:100-149) or HemisphericCamera::GenerateRay
//              (cameras/hemispheric.cpp:15-41)
//   cast       one persistent launch over three ray queues (traverse.cuh, per-ray query kind):
//              Scene::Intersect -> BVHAccel::Intersect for the paths' next rays, and the two
//              rays the previous shade left behind for EstimateDirect —
//                shadow  VisibilityTester::Unoccluded -> BVHAccel::IntersectP (any hit)
//                mis     the BSDF-sampled ray: closest hit, Le credited when the first hit is the
//                        sampled light (integrator.cpp:196-211)
//              One launch instead of three: one tail, and late bounces still fill the machine.
//   shade      emission, Material::ComputeScatteringFunctions, UniformSampleOneLight /
//              EstimateDirect's sampling half (integrator.cpp:85-215), BSDF::Sample_f, beta
//              update, Russian roulette (path.cpp:88-190).  Emits at most one shadow ray, one
//              MIS ray and one continuation ray per path.
//   film       the NaN / negative / infinite scrub (integrator.cpp:293-314) and
//              FilmTile::AddSample (film.h:153-193) into an {R,G,B,weight} film
//
// Queues are compacted with warp-aggregated atomics (ballot + popc prefix); every kernel is a
// persistent grid-stride loop over a device-side counter, so a whole render is enqueued
// without host synchronisation.  Random numbers: Sampler "random" = one PCG32 stream per
// (pixel, sample), sequence = sample << 32 | pixel; Sampler "halton" = the reference's own
// Halton points (halton.cuh); either way drawn in exactly the order Li draws them.
#include "shading.cuh"
#include "traverse.cuh"
#include "wavefront.h"
#include "camera.cuh"
#include "texture.cuh"

namespace iile {

namespace {

constexpr int kBlock = 128;

// Path state and queue items are read once and written once per kernel; IILE_STREAM_STATE=1 marks those
// accesses evict-first (ld.global.cs / st.global.cs) so that the gigabytes of state streaming through the L2
// would not displace the BVH nodes and triangles the traversal launches of the other lanes are reading
// (cfg5: 0.9 GB of nodes and triangles against a 126 MB L2).  Measured on B200 it changes nothing (cfg5
// 711 -> 708, cfg3 244 -> 243, cfg1 437 -> 435 Msamples/s, profiles/r8j_ab_stream_state.txt): off.
#ifndef IILE_STREAM_STATE
#define IILE_STREAM_STATE 0
#endif
#if IILE_STREAM_STATE
#define LD_STATE(p) __ldcs(p)
#define ST_STATE(p, v) __stcs((p), (v))
#else
#define LD_STATE(p) (*(p))
#define ST_STATE(p, v) (*(p) = (v))
#endif

// ---- warp-aggregated queue push ---------------------------------------------------
__device__ __forceinline__ uint32_t queueSlot(uint32_t *counter, bool want) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);  // every lane of the warp calls this
    if (!want) return 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1));
}

__device__ __forceinline__ V3f f4(float4 v) { return V3f(v.x, v.y, v.z); }
__device__ __forceinline__ float4 mk4(V3f v, float w) { return make_float4(v.x, v.y, v.z, w); }

// ---- generate -----------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) generateKernel(RenderConsts rc, PathState st, Queues q,
                                                         uint64_t firstWork, uint32_t nWork,
                                                         HemiOutputs hemi) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ((nWork + 31) & ~31u);
         i += gridDim.x * blockDim.x) {
        bool valid = i < nWork;
        int px = 0, py = 0;
        uint32_t sIdx = 0, pixelIdx = 0;
        const int sbw = rc.sampleBounds[2] - rc.sampleBounds[0];
        if (valid) {
            const uint64_t w = firstWork + i;
            const uint32_t sppLocal = (uint32_t)(rc.sampleEnd - rc.sampleBegin);
            if (rc.hemiMode) {
                // work = ((hemisphere * size*size) + pixel) * spp + s ; pixels of [0,size)^2 only
                const uint32_t s = (uint32_t)(w % sppLocal);
                const uint64_t r = w / sppLocal;
                const uint32_t pix = (uint32_t)(r % (uint32_t)(rc.hemiSize * rc.hemiSize));
                const uint32_t h = (uint32_t)(r / (uint32_t)(rc.hemiSize * rc.hemiSize));
                px = pix % rc.hemiSize;
                py = pix / rc.hemiSize;
                sIdx = h * sppLocal + s;
            } else {
                // 16x16 tiles in Render's order; inside a tile pixels row-major, samples innermost
                const uint32_t perTile = 256u * sppLocal;
                const uint32_t tile = rc.tileBegin + (uint32_t)(w / perTile);
                const uint32_t r = (uint32_t)(w % perTile);
                const uint32_t pit = r / sppLocal;
                sIdx = rc.sampleBegin + r % sppLocal;
                px = rc.sampleBounds[0] + (tile % rc.tilesX) * 16 + (pit & 15);
                py = rc.sampleBounds[1] + (tile / rc.tilesX) * 16 + (pit >> 4);
                valid = px < rc.sampleBounds[2] && py < rc.sampleBounds[3] && px >= rc.pixelBounds[0] &&
                        px < rc.pixelBounds[2] && py >= rc.pixelBounds[1] && py < rc.pixelBounds[3];
            }
            pixelIdx = (uint32_t)(py - rc.sampleBounds[1]) * sbw + (uint32_t)(px - rc.sampleBounds[0]);
        }
        const uint32_t slot = queueSlot(&q.counters[0], valid);
        if (!valid) continue;
        // GetCameraSample (sampler.cpp:46-52): pFilm (2), time (1), pLens (2)
        SampleStream g;
        g.tables = rc.halton.tables;
        g.sobol = rc.halton.sobol.matrices;
        g.dim = 0;
        float fx, fy, lu = 0, lv = 0;
        if (g.tables) {
            // HaltonSampler::SampleDimension (halton.cpp:115-122): the first two dimensions are
            // unscrambled and pre-divided so that they enumerate the pixel's area
            const uint64_t index = haltonIndexForSample(rc.halton, px, py, sIdx);
            g.pcg.state = index;
            g.pcg.inc = 0;
            const float u0 = rc.halton.samplePixelCenter ? 0.5f : radicalInverseBase2(index >> rc.halton.baseExponents[0]);
            const float u1 = rc.halton.samplePixelCenter ? 0.5f : radicalInverse(3, index / (uint64_t)rc.halton.baseScales[1]);
            fx = (float)px + u0;
            fy = (float)py + u1;
            if (rc.lensRadius > 0) {
                lu = haltonScrambled(g.tables, index, 3);
                lv = haltonScrambled(g.tables, index, 4);
            }
            g.dim = rc.firstDimension;
        } else if (g.sobol) {
            // SobolSampler::GetIndexForSample / SampleDimension (sobol.cpp:41-59): the first two dimensions
            // are scaled to the 2^m grid and taken relative to the pixel
            const SobolConsts &sb = rc.halton.sobol;
            const uint64_t index = sobolIntervalToIndex(sb, sIdx, (uint32_t)(px - rc.sampleBounds[0]),
                                                        (uint32_t)(py - rc.sampleBounds[1]));
            g.pcg.state = index;
            g.pcg.inc = 0;
            float s0 = sobolSample(g.sobol, index, 0) * (float)sb.resolution + (float)rc.sampleBounds[0];
            float s1 = sobolSample(g.sobol, index, 1) * (float)sb.resolution + (float)rc.sampleBounds[1];
            s0 = clampf(s0 - (float)px, 0.f, IILE_ONE_MINUS_EPS);
            s1 = clampf(s1 - (float)py, 0.f, IILE_ONE_MINUS_EPS);
            fx = (float)px + s0;
            fy = (float)py + s1;
            if (rc.lensRadius > 0) {
                lu = sobolSample(g.sobol, index, 3);
                lv = sobolSample(g.sobol, index, 4);
            }
            g.dim = rc.firstDimension;
        } else {
            g.pcg.setSequence(((uint64_t)sIdx << 32) | pixelIdx);
            fx = (float)px + g.pcg.uniformFloat();
            fy = (float)py + g.pcg.uniformFloat();
            (void)g.pcg.uniformFloat();
            lu = g.pcg.uniformFloat();
            lv = g.pcg.uniformFloat();
        }
        V3f o(0, 0, 0), d;
        const float *c2w = rc.cameraToWorld;
        if (rc.hemiMode) {
            const uint32_t h = sIdx / (uint32_t)(rc.sampleEnd - rc.sampleBegin);
            c2w = hemi.cams + 32 * (size_t)h;  // per-hemisphere {cameraToWorld, worldToCamera}
            d = hemiCameraSpaceDir(rc.hemiSize, fx, fy);
        } else
            cameraSpaceRay(rc, fx, fy, lu, lv, &o, &d);
        rayToWorld(c2w, &o, &d);
        st.pixel[slot] = pixelIdx;
        st.sample[slot] = sIdx;
        st.rng[slot] = g.pack();
        ST_STATE(st.rayO + slot, mk4(o, fx));
        ST_STATE(st.rayD + slot, mk4(d, fy));
        ST_STATE(st.beta + slot, make_float4(1, 1, 1, 1));
        ST_STATE(st.L + slot, make_float4(0, 0, 0, __int_as_float(0)));
        ST_STATE(st.pendS + slot, make_float4(0, 0, 0, 0));
        ST_STATE(st.pendM + slot, make_float4(0, 0, 0, 0));
        if (st.texRecord) st.texRecord[8 * (size_t)slot + 7].w = 0.f;  // the ray is the camera's (see textureKernel)
        q.active[0][slot] = slot;
        atomicAdd(&q.stats[2], 1ull);
    }
}

// The last block to finish puts the work counter back to zero for the next launch.
__device__ __forceinline__ void releaseWorkCounter(uint32_t *counters) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicInc(&counters[1], gridDim.x - 1) == gridDim.x - 1) counters[0] = 0;
    }
}

// ---- classify: partition the active queue by shading class ---------------------------------------
// Each block takes segments of kClassifySeg consecutive queue entries, ranks every entry among
// the entries of its class inside the segment (ballot / popc per warp, a short scan over the
// warps), reserves room in the class queues with one atomicAdd per class, and scatters.  The
// shade kernel then sweeps the class queues back to back: a warp — and, since the grid walks the
// queues in order, every SM at any moment — runs one material's code (the kernel is ~12k
// instructions; mixed classes thrash the instruction cache: profiles/r1g, r1l).  Inside a segment
// the queue order (neighbouring pixels, similar rays) is preserved.
constexpr int kClassifySeg = 1024;

__global__ void __launch_bounds__(kClassifySeg) classifyKernel(DeviceScene sc, PathState st, Queues q, int cur) {
    __shared__ uint32_t warpCount[kNumClasses][kClassifySeg / 32];
    __shared__ uint32_t classBase[kNumClasses];
    const uint32_t n = q.counters[cur];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t seg = blockIdx.x; (uint64_t)seg * kClassifySeg < n; seg += gridDim.x) {
        const uint32_t i = seg * kClassifySeg + threadIdx.x;
        const bool live = i < n;
        uint32_t slot = 0;
        int cls = -1;
        if (live) {
            slot = q.active[cur][i];
            const int leafSlot = st.hit[slot].x;
            cls = leafSlot < 0 ? kClassMiss
                               : (int)((__float_as_uint(__ldg(sc.leafPrims + 3 * leafSlot + 1).w) >> 16) & 0xfu);
        }
        uint32_t rankInWarp = 0;
#pragma unroll
        for (int c = 0; c < kNumClasses; ++c) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == c);
            if (cls == c) rankInWarp = __popc(m & ((1u << lane) - 1));
            if (lane == 0) warpCount[c][warp] = __popc(m);
        }
        __syncthreads();
        if (threadIdx.x < kNumClasses) {
            uint32_t acc = 0;
            for (int w = 0; w < kClassifySeg / 32; ++w) {
                const uint32_t v = warpCount[threadIdx.x][w];
                warpCount[threadIdx.x][w] = acc;
                acc += v;
            }
            // room for this segment's paths of the class in the class queue
            classBase[threadIdx.x] = acc ? atomicAdd(&q.counters[kClassCounter0 + threadIdx.x], acc) : 0u;
        }
        __syncthreads();
        if (live) q.byClass[cls][classBase[cls] + warpCount[cls][warp] + rankInWarp] = slot;
        __syncthreads();
    }
}

// Re-derive the SurfaceInteraction of the hit found by intersectKernel (the winning
// primitive's test is repeated: same inputs, same bits).
static __device__ __noinline__ bool surfaceAt(const DeviceScene &sc, int leafSlot, V3f o, V3f d, Surface *s,
                                          float *tOut) {
    const float4 v0 = __ldg(sc.leafPrims + 3 * leafSlot);
    const int primIdx = __float_as_int(v0.w);
    const IilePrim pr = sc.prims[primIdx];
    RayCtx r;
    initRay(r, o.x, o.y, o.z, d.x, d.y, d.z, INFINITY);
    if (pr.type == IILE_PRIM_TRIANGLE) {
        const float4 v1 = __ldg(sc.leafPrims + 3 * leafSlot + 1), v2 = __ldg(sc.leafPrims + 3 * leafSlot + 2);
        float t, b0, b1, b2;
        if (!triangleTest(r, INFINITY, make_float3(v0.x, v0.y, v0.z), make_float3(v1.x, v1.y, v1.z),
                          make_float3(v2.x, v2.y, v2.z), &t, &b0, &b1, &b2))
            return false;
        triangleSurface(sc, pr, b0, b1, b2, d, s);
        *tOut = t;
    } else {
        const IileSphere &sp = sc.spheres[pr.v[0]];
        SphereHit sh;
        float t;
        if (!sphereQuadric(sp, o.x, o.y, o.z, d.x, d.y, d.z, INFINITY, &t, &sh)) return false;
        sphereSurface(sp, pr, sh, s);
        *tOut = t;
    }
    return true;
}

// AreaLight::L, diffuse.h:56-58
__device__ __forceinline__ V3f emitted(const DeviceScene &sc, int light, V3f n, V3f w) {
    if (light < 0) return V3f(0, 0, 0);
    const IileLight &L = sc.lights[light];
    return (IILE_LIGHT_TWOSIDED(L) || dot(n, w) > 0) ? ld3(L.L) : V3f(0, 0, 0);
}

// ---- shade ------------------------------------------------------------------------------
// The body is ~12k instructions (≈190 KB of SASS, more than the instruction caches hold) and every
// path runs most of it once, so warps that drift apart each stream the whole kernel through
// the i-cache on their own: profiles/r1q showed 58 % of the stall samples as "no instruction".
// The body is therefore cut into phases separated by block barriers (IILE_SHADE_PHASES): the warps
// of a block — one large block per SM — walk the code together and share every fetched line.
// The arithmetic and the order of the random draws of one path are unchanged.
#ifndef IILE_SHADE_BLOCK
#define IILE_SHADE_BLOCK 512
#endif
#ifndef IILE_SHADE_PHASES
#define IILE_SHADE_PHASES 1
#endif
#ifndef IILE_SHADE_BLOCK_PUSH
#define IILE_SHADE_BLOCK_PUSH 1
#endif
constexpr int kShadeBlock = IILE_SHADE_BLOCK;

// MATERIAL < 0: the generic kernel — every material and light through the out-of-line functions of
// shading.cuh, the shading classes of `classMask` back to back, phase barriers (see above).
// MATERIAL = IILE_MAT_MATTE / IILE_MAT_PLASTIC: that one class with its BSDF built and evaluated in line
// (one or two lobes, held in registers), the triangle surface in line (quadrics are classed kClassOther),
// and — SIMPLE: every light of the scene is a point light or a diffuse area light on a triangle — the
// light sample in line as well.  The same expressions in the same order as the generic kernel; what
// changes is that nothing is passed through local memory and the kernel needs half the registers, so
// twice the warps hide the latency of the dependent loads (profiles/r7b_shade_cfg5_ncu_summary.csv:
// the generic kernel sits at 26 % issue-active with long-scoreboard and barrier stalls on top).
constexpr int kFastBlock = 256;
#ifndef IILE_FAST_MINBLOCKS
#define IILE_FAST_MINBLOCKS 2
#endif
#ifndef IILE_FAST_EARLY_PUSH
#define IILE_FAST_EARLY_PUSH 1
#endif
template <int MATERIAL, bool SIMPLE>
__global__ void __launch_bounds__(MATERIAL < 0 ? kShadeBlock : kFastBlock, MATERIAL < 0 ? 1 : IILE_FAST_MINBLOCKS)
    shadeKernelT(DeviceScene sc, RenderConsts rc, PathState st, Queues q, int cur, HemiOutputs hemi, int classMask) {
    constexpr bool kGeneric = MATERIAL < 0;
    constexpr int kThreads = kGeneric ? kShadeBlock : kFastBlock;
    constexpr int KINDS = MATERIAL == IILE_MAT_MATTE ? kMatteLobes : (MATERIAL == IILE_MAT_PLASTIC ? kPlasticLobes : kAllLobes);
    constexpr int MAXN = MATERIAL == IILE_MAT_MATTE ? 1 : (MATERIAL == IILE_MAT_PLASTIC ? 2 : kMaxLobes);
#undef SHADE_PHASE_BARRIER
#if IILE_SHADE_PHASES
#define SHADE_PHASE_BARRIER() \
    if (kGeneric) __syncthreads()
#else
#define SHADE_PHASE_BARRIER()
#endif
    const int nxt = cur ^ 1;
#if IILE_SHADE_BLOCK_PUSH
    __shared__ uint32_t pushCount[4][kThreads / 32];
    __shared__ uint32_t pushBase[4];
#endif
    // the class queues of this launch back to back: [0, end[0]) misses, [end[0], end[1]) matte, ...
    uint32_t end[kNumClasses];
    {
        uint32_t acc = 0;
        for (int c = 0; c < kNumClasses; ++c) {
            if ((classMask >> c) & 1) acc += q.counters[kClassCounter0 + c];
            end[c] = acc;
        }
    }
    const uint32_t n = end[kNumClasses - 1];
    // block-uniform trip count: every thread of a block reaches every barrier
    for (uint32_t base = blockIdx.x * kThreads; base < n; base += gridDim.x * kThreads) {
        const uint32_t i = base + threadIdx.x;
        const bool live = i < n;
        bool wantShadow = false, wantMis = false, wantNext = false, wantDone = false;
        uint32_t slot = 0;
        ShadowItem sh;
        MisItem mi;
        // ---- phase A: path state, SurfaceInteraction, emission ------------------------------------
        float4 o4, d4;
        V3f o, d, beta, L, newPendS(0, 0, 0);
        float etaScale = 1;
        int bounces = 0;
        bool specularBounce = false, found = false, shadeIt = false;
        SampleStream g;
        Surface s;
        if (live) {
            int cls = 0;
            while (i >= end[cls]) ++cls;
            slot = q.byClass[cls][i - (cls ? end[cls - 1] : 0u)];
            o4 = LD_STATE(st.rayO + slot);
            d4 = LD_STATE(st.rayD + slot);
            const float4 beta4 = LD_STATE(st.beta + slot), L4 = LD_STATE(st.L + slot);
            const float4 pS = LD_STATE(st.pendS + slot), pM = LD_STATE(st.pendM + slot);
            const int2 hit = LD_STATE(st.hit + slot);
            o = f4(o4), d = f4(d4), beta = f4(beta4);
            // direct light of the previous vertex, resolved by the cast launch in between
            L = f4(L4) + (f4(pS) + f4(pM));
            etaScale = beta4.w;
            const int flags = __float_as_int(L4.w);
            bounces = flags & 0xffff;
            specularBounce = (flags >> 16) & 1;
            g.unpack(st.rng[slot], st.sample[slot], st.pixel[slot], rc.halton);

            float tHit = 0;
            if constexpr (kGeneric)
                found = hit.x >= 0 && surfaceAt(sc, hit.x, o, d, &s, &tHit);
            else {  // a triangle with this kernel's material: surfaceAt's triangle branch, in line
                const float4 v0 = __ldg(sc.leafPrims + 3 * hit.x), v1 = __ldg(sc.leafPrims + 3 * hit.x + 1),
                             v2 = __ldg(sc.leafPrims + 3 * hit.x + 2);
                const IilePrim pr = sc.prims[__float_as_int(v0.w)];
                RayCtx r;
                initRay(r, o.x, o.y, o.z, d.x, d.y, d.z, INFINITY);
                float b0, b1, b2;
                found = triangleTest(r, INFINITY, make_float3(v0.x, v0.y, v0.z), make_float3(v1.x, v1.y, v1.z),
                                     make_float3(v2.x, v2.y, v2.z), &tHit, &b0, &b1, &b2);
                if (found) triangleSurfaceInl(sc, pr, b0, b1, b2, d, &s);
            }
            // IISPTdIntegrator: first-hit distance and camera-space normal (iispt_d.cpp:99-115)
            if (rc.hemiMode && bounces == 0 && hemi.distance) {
                const uint32_t sppLocal = (uint32_t)(rc.sampleEnd - rc.sampleBegin);
                const uint32_t sIdx = st.sample[slot];
                if (sIdx % sppLocal == sppLocal - 1) {  // set_camera_coord overwrites: the pixel's last sample stays
                    const size_t h = sIdx / sppLocal;
                    const size_t at = h * rc.hemiSize * rc.hemiSize + st.pixel[slot];
                    if (found) {
                        V3f cv = s.p - o;
                        hemi.distance[at] = sqrtf(dot(cv, cv));
                        const float *mi2 = hemi.cams + 32 * h + 16;  // WorldToCamera(Normal3f) multiplies by mInv transposed
                        V3f nc(mi2[0] * s.n.x + mi2[4] * s.n.y + mi2[8] * s.n.z,
                               mi2[1] * s.n.x + mi2[5] * s.n.y + mi2[9] * s.n.z,
                               mi2[2] * s.n.x + mi2[6] * s.n.y + mi2[10] * s.n.z);
                        hemi.normals[3 * at] = nc.x;
                        hemi.normals[3 * at + 1] = nc.y;
                        hemi.normals[3 * at + 2] = nc.z;
                    } else {
                        hemi.distance[at] = -1.f;
                        hemi.normals[3 * at] = hemi.normals[3 * at + 1] = hemi.normals[3 * at + 2] = 0.f;
                    }
                }
            }
            // path.cpp:91-101 (no infinite lights in scope); the d-integrator skips Le at bounce 0
            if (found && (bounces == 0 || specularBounce) && !(rc.hemiMode && bounces == 0))
                L = L + beta * emitted(sc, s.light, s.n, -d);
            // escaped ray: scene.infiniteLights' Le (path.cpp:96-99; the d-integrator only past bounce 0,
            // iispt_d.cpp:124-130)
            if (!found && sc.nInfinite > 0 && (bounces == 0 || specularBounce) && !(rc.hemiMode && bounces == 0))
                for (uint32_t k = 0; k < sc.nLights; ++k)
                    if (sc.lights[k].type == IILE_LIGHT_INFINITE) L = L + beta * envLe(sc.lights[k], d);
            const bool alive = found && bounces < rc.maxDepth;  // path.cpp:104
            if (alive && s.material < 0) {
                // interface-only primitive: continue through it without counting a bounce (path.cpp:111-116)
                V3f no = offsetRayOrigin(s.p, s.pError, s.n, d);
                st.rayO[slot] = mk4(no, o4.w);
                // isect.SpawnRay(ray.d) is a plain Ray: the camera's differentials end here (texture stage)
                if (st.texRecord) st.texRecord[8 * (size_t)slot + 7].w = 1.f;
                wantNext = true;
            } else if (alive)
                shadeIt = true;
            else
                wantDone = true;
        }
        SHADE_PHASE_BARRIER();
        // ---- phase B: BSDF, UniformSampleOneLight's light pick and light sample -------------------
        Bsdf bsdf;
        LightSample ls;
        bool direct = false, delta = false;
        int lightNum = 0;
        float pickPdf = 0, us0 = 0, us1 = 0;
        constexpr int nonSpec = BSDF_ALL & ~BSDF_SPECULAR;
        if (shadeIt) {
            if constexpr (kGeneric) {
                const IileMaterial *mp = &sc.materials[s.material];
                if (mp->texSet) {
                    // image textures / bump map: textureKernel evaluated them for this hit (record in HBM: the
                    // resolved material, then the bump-mapped shading normal and tangent)
                    const float4 *rec = st.texRecord + 8 * (size_t)slot;
                    mp = reinterpret_cast<const IileMaterial *>(rec);
                    s.ns = f4(rec[6]);
                    s.ss = f4(rec[7]);
                    s.ts = cross(s.ns, s.ss);
                }
                makeBsdf(*mp, s, &bsdf);
            } else makeBsdfFast<MATERIAL>(sc.materials[s.material], s, &bsdf);
            int nonSpecLobes = 0;
            if constexpr (kGeneric) nonSpecLobes = bsdf.numComponents(nonSpec);
            else {
#pragma unroll
                for (int k = 0; k < MAXN; ++k)
                    if (k < bsdf.n && (bsdf.lobe[k].type & nonSpec) == bsdf.lobe[k].type) ++nonSpecLobes;
            }
            // UniformSampleOneLight (integrator.cpp:85-106)
            if (nonSpecLobes > 0 && sc.nLights > 0) {
                const float *table = lightTableAt(sc, s.p);
                lightNum = sampleDiscrete(table, (int)sc.nLights, g.uniformFloat(), &pickPdf);
                if (pickPdf != 0) {
                    const float ul0 = g.uniformFloat(), ul1 = g.uniformFloat();
                    us0 = g.uniformFloat();
                    us1 = g.uniformFloat();
                    const IileLight &light = sc.lights[lightNum];
                    // IsDeltaLight: point, spot, distant
                    delta = light.type != IILE_LIGHT_DIFFUSE_AREA && light.type != IILE_LIGHT_INFINITE;
                    direct = true;
                    // light-sampling half of EstimateDirect (integrator.cpp:118-165)
                    if constexpr (!kGeneric && SIMPLE) sampleLightSimple(sc, light, s.p, s.pError, s.n, ul0, ul1, &ls);
                    else sampleLight(sc, light, s.p, s.pError, s.n, ul0, ul1, &ls);
                }
            }
        }
        SHADE_PHASE_BARRIER();
        // ---- phase C: BSDF value and pdf towards the light sample -> shadow ray -------------------
        if (direct && ls.pdf > 0 && !isBlack(ls.Li)) {
            V3f f = bsdfFT<KINDS, MAXN, kGeneric>(bsdf, s.wo, ls.wi, nonSpec) * absDot(ls.wi, s.ns);
            float scatteringPdf = bsdfPdfT<KINDS, MAXN, kGeneric>(bsdf, s.wo, ls.wi, nonSpec);
            if (!isBlack(f)) {
                V3f Ld;
                if (delta) Ld = sdiv(f * ls.Li, ls.pdf);
                else Ld = sdiv(f * ls.Li * powerHeuristic(ls.pdf, scatteringPdf), ls.pdf);
                // SpawnRayTo (interaction.h:68-78): unnormalised d, tMax = 1 - ShadowEpsilon
                const V3f so = ls.pOrigin;
                const V3f sd = ls.pTarget - so;
                sh.o = mk4(so, __int_as_float((int)slot));
                sh.d = mk4(sd, 1 - IILE_SHADOW_EPS);
                newPendS = beta * sdiv(Ld, pickPdf);
                wantShadow = true;
            }
        }
        if constexpr (!kGeneric && IILE_FAST_EARLY_PUSH) {
            // the item leaves now (one atomic per warp) instead of riding along to the end of the
            // iteration: eight registers less across the BSDF sampling below
            const uint32_t k = queueSlot(&q.counters[2], wantShadow);
            if (wantShadow) {
                ST_STATE(&q.shadow[k].o, sh.o);
                ST_STATE(&q.shadow[k].d, sh.d);
            }
            wantShadow = false;
        }
        SHADE_PHASE_BARRIER();
        // ---- phase D: BSDF-sampling half (integrator.cpp:167-213) -> MIS ray ----------------------
        if (direct && !delta) {
            V3f wi;
            float scatteringPdf;
            int sampledType;
            V3f f = bsdfSampleT<KINDS, MAXN, kGeneric>(bsdf, s.wo, &wi, us0, us1, &scatteringPdf, nonSpec, &sampledType);
            if (scatteringPdf > 0) f = f * absDot(wi, s.ns);
            if (!isBlack(f) && scatteringPdf > 0) {
                V3f ro = offsetRayOrigin(s.p, s.pError, s.n, wi);
                const IileLight &light = sc.lights[lightNum];
                float lightPdf;
                if constexpr (!kGeneric && SIMPLE) lightPdf = areaLightPdfTriangle(sc, light, s.p, ro, wi);
                else
                    lightPdf = light.type == IILE_LIGHT_INFINITE ? envPdf(light, wi)
                                                                 : areaLightPdf(sc, light, s.p, s.pError, s.n, ro, wi);
                if (lightPdf != 0) {
                    float weight = powerHeuristic(scatteringPdf, lightPdf);
                    mi.o = mk4(ro, __int_as_float((int)slot));
                    mi.d = mk4(wi, __int_as_float(lightNum));
                    mi.contrib = mk4(beta * sdiv(sdiv(f * weight, scatteringPdf), pickPdf), __int_as_float(-1));
                    wantMis = true;
                }
            }
        }
        if constexpr (!kGeneric && IILE_FAST_EARLY_PUSH) {
            const uint32_t k = queueSlot(&q.counters[3], wantMis);
            if (wantMis) {
                ST_STATE(&q.mis[k].o, mi.o);
                ST_STATE(&q.mis[k].d, mi.d);
                ST_STATE(&q.mis[k].contrib, mi.contrib);
            }
            wantMis = false;
        }
        SHADE_PHASE_BARRIER();
        // ---- phase E: continuation (path.cpp:133-156) and Russian roulette ------------------------
        if (shadeIt) {
            // `Vector3f wo = -ray.d` (path.cpp:132, iispt_d.cpp likewise): NOT the interaction's normalised wo that the
            // direct-lighting estimate uses — a bounce ray's direction is a unit vector only to an ulp, and the
            // reflected / refracted direction inherits the difference
            const V3f wo = -d;
            const float u0 = g.uniformFloat(), u1 = g.uniformFloat();
            V3f wi;
            float pdf;
            int sampled;
            V3f f = bsdfSampleT<KINDS, MAXN, kGeneric>(bsdf, wo, &wi, u0, u1, &pdf, BSDF_ALL, &sampled);
            bool cont = !(isBlack(f) || pdf == 0.f);
            if (cont) {
                beta = beta * sdiv(f * absDot(wi, s.ns), pdf);
                float by = luminance(beta);
                if (by < 0.f || isnan(by)) cont = false;  // IILE addition, path.cpp:143-145
            }
            if (cont) {
                specularBounce = (sampled & BSDF_SPECULAR) != 0;
                if ((sampled & BSDF_SPECULAR) && (sampled & BSDF_TRANSMISSION)) {
                    float eta = bsdf.eta;
                    etaScale *= (dot(wo, s.n) > 0) ? (eta * eta) : 1 / (eta * eta);
                }
                V3f no = offsetRayOrigin(s.p, s.pError, s.n, wi);
                // Russian roulette, path.cpp:182-190
                V3f rr = beta * etaScale;
                float mx = fmaxf(rr.x, fmaxf(rr.y, rr.z));
                if (mx < rc.rrThreshold && bounces > 3) {
                    float qq = fmaxf(.05f, 1 - mx);
                    if (g.uniformFloat() < qq) cont = false;
                    else beta = sdiv(beta, 1 - qq);
                }
                if (cont) {
                    ST_STATE(st.rayO + slot, mk4(no, o4.w));
                    ST_STATE(st.rayD + slot, mk4(wi, d4.w));
                    ++bounces;
                    wantNext = true;
                }
            }
            if (!wantNext) wantDone = true;
        }
        SHADE_PHASE_BARRIER();
        // ---- phase F: write back, queue pushes ------------------------------------------------------
        if (live) {
            ST_STATE(st.beta + slot, mk4(beta, etaScale));
            ST_STATE(st.L + slot, mk4(L, __int_as_float((bounces & 0xffff) | (specularBounce ? 1 << 16 : 0))));
            st.rng[slot] = g.pack();
            ST_STATE(st.pendS + slot, mk4(newPendS, 0));
            ST_STATE(st.pendM + slot, make_float4(0, 0, 0, 0));
        }
#if IILE_SHADE_BLOCK_PUSH
        // Block-aggregated pushes: the four queues' per-warp counts meet in shared memory, four
        // threads reserve the block's room with one atomicAdd each — concurrently, one global
        // round trip per block iteration instead of four dependent ones per warp (profiles/r1q:
        // 16 % of the kernel's stall samples sat on the per-warp atomics' return values).
        {
            const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const unsigned lt = (1u << lane) - 1;
            const bool want[4] = {wantNext, wantShadow, wantMis, wantDone};
            uint32_t rank[4];
#pragma unroll
            for (int qi = 0; qi < 4; ++qi) {
                const unsigned m = __ballot_sync(0xffffffffu, want[qi]);
                rank[qi] = __popc(m & lt);
                if (lane == 0) pushCount[qi][warp] = __popc(m);
            }
            __syncthreads();
            if (threadIdx.x < 4) {
                uint32_t acc = 0;
                for (int w = 0; w < kThreads / 32; ++w) {
                    const uint32_t v = pushCount[threadIdx.x][w];
                    pushCount[threadIdx.x][w] = acc;
                    acc += v;
                }
                const int counter = threadIdx.x == 0 ? nxt : (threadIdx.x == 1 ? 2 : (threadIdx.x == 2 ? 3 : 4));
                pushBase[threadIdx.x] = acc ? atomicAdd(&q.counters[counter], acc) : 0u;
            }
            __syncthreads();
            if (wantNext) q.active[nxt][pushBase[0] + pushCount[0][warp] + rank[0]] = slot;
            if (wantShadow) q.shadow[pushBase[1] + pushCount[1][warp] + rank[1]] = sh;
            if (wantMis) q.mis[pushBase[2] + pushCount[2][warp] + rank[2]] = mi;
            if (wantDone) q.done[pushBase[3] + pushCount[3][warp] + rank[3]] = slot;
            __syncthreads();  // pushCount / pushBase are rewritten by the next iteration
        }
#else
        uint32_t k = queueSlot(&q.counters[nxt], wantNext);
        if (wantNext) q.active[nxt][k] = slot;
        k = queueSlot(&q.counters[2], wantShadow);
        if (wantShadow) q.shadow[k] = sh;
        k = queueSlot(&q.counters[3], wantMis);
        if (wantMis) q.mis[k] = mi;
        k = queueSlot(&q.counters[4], wantDone);
        if (wantDone) q.done[k] = slot;
#endif
    }
}

// The texture stage of a bounce (scenes with image textures or bump maps only): for every path of class
// kClassOther whose hit has a textured material, the material's textures are evaluated for this hit and the
// bump map applied (texture.cuh), and the result — the material record with every texture replaced by its
// value, the shading normal and tangent — is left in HBM for the shade kernel, which then builds the BSDF
// exactly as for a constant material.  A stage of its own keeps the texture filter's registers and call
// depth out of the shade kernel.
__global__ void __launch_bounds__(kBlock) textureKernel(DeviceScene sc, RenderConsts rc, PathState st, Queues q,
                                                        HemiOutputs hemi) {
    const uint32_t n = q.counters[kClassCounter0 + kClassOther];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = q.byClass[kClassOther][i];
        const int2 hit = st.hit[slot];
        if (hit.x < 0) continue;
        const int primIdx = __float_as_int(__ldg(sc.leafPrims + 3 * hit.x).w);
        const IilePrim pr = sc.prims[primIdx];
        if (pr.material < 0 || !sc.materials[pr.material].texSet) continue;
        const float4 o4 = st.rayO[slot], d4 = st.rayD[slot];
        Surface s;
        float tHit;
        if (!surfaceAt(sc, hit.x, f4(o4), f4(d4), &s, &tHit)) continue;
        const int bounces = __float_as_int(st.L[slot].w) & 0xffff;
        IileMaterial m = sc.materials[pr.material];
        // RenderView's camera rays carry the HemisphericCamera's differentials (iispt_d.cpp:403-407)
        const float *hemiCam = nullptr;
        if (rc.hemiMode) hemiCam = hemi.cams + 32 * (size_t)(st.sample[slot] / (uint32_t)(rc.sampleEnd - rc.sampleBegin));
        // differentials: the camera ray's only — bounce 0, and no interface-only primitive crossed on the way
        // (path.cpp:111-116 continues with a plain Ray; the shade kernel leaves a mark in the record's last word)
        float4 *rec = st.texRecord + 8 * (size_t)slot;
        const float crossed = rec[7].w;
        const bool cameraRay = bounces == 0 && crossed == 0.f;
        V3f lensPt(0, 0, 0);
        const bool lens = cameraRay && !hemiCam && rc.lensRadius > 0;
        if (lens) lensPt = lensPointOf(rc, st.pixel[slot], st.sample[slot]);
        resolveTextured(sc, rc, primIdx, f4(o4), f4(d4), cameraRay, o4.w, d4.w, hemiCam, 1, &s, &m, lens ? &lensPt : nullptr);
        const float4 *src = reinterpret_cast<const float4 *>(&m);
#pragma unroll
        for (int k = 0; k < 6; ++k) rec[k] = src[k];
        rec[6] = mk4(s.ns, 0);
        rec[7] = mk4(s.ss, crossed);
    }
}

// The texture stage of the direct integrators (scenes with textures only): every hit about to be shaded gets its
// record — material with the textures evaluated, shading frame, and what SpecularReflect / SpecularTransmit need to
// hand the ray's differentials on (PathState::dirRec).  The camera ray's differentials are re-derived from pFilm;
// a spawned ray's are in PathState::rayDiff.
__global__ void __launch_bounds__(kBlock) textureDirectKernel(DeviceScene sc, RenderConsts rc, PathState st, Queues q, int cur) {
    const uint32_t n = q.counters[cur];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = q.active[cur][i];
        const int2 hit = st.hit[slot];
        if (hit.x < 0) continue;
        const int primIdx = __float_as_int(__ldg(sc.leafPrims + 3 * hit.x).w);
        const IilePrim pr = sc.prims[primIdx];
        if (pr.material < 0) continue;
        const float4 o4 = st.rayO[slot], d4 = st.rayD[slot];
        const V3f o = f4(o4), d = f4(d4);
        Surface s;
        float tHit;
        if (!surfaceAt(sc, hit.x, o, d, &s, &tHit)) continue;
        IileMaterial m = sc.materials[pr.material];
        bool has = false;
        V3f rxO(0, 0, 0), rxD(0, 0, 0), ryO(0, 0, 0), ryD(0, 0, 0);
        if (!((__float_as_int(st.L[slot].w) >> 28) & 1)) {
            V3f lensPt(0, 0, 0);
            if (rc.lensRadius > 0) lensPt = lensPointOf(rc, st.pixel[slot], st.sample[slot]);
            cameraDifferentials(rc, nullptr, o4.w, d4.w, o, d, &rxO, &rxD, &ryO, &ryD, rc.lensRadius > 0 ? &lensPt : nullptr);
            has = true;
        } else {
            const float4 *rdf = st.rayDiff + 4 * (size_t)slot;
            const float4 a = rdf[0];
            if (a.w != 0.f) {
                has = true;
                rxO = f4(a), ryO = f4(rdf[1]), rxD = f4(rdf[2]), ryD = f4(rdf[3]);
            }
        }
        V3f dpdx(0, 0, 0), dpdy(0, 0, 0), dndx(0, 0, 0), dndy(0, 0, 0);
        resolveDirectVertex(sc, primIdx, o, d, has, rxO, rxD, ryO, ryD, &s, &m, &dpdx, &dpdy, &dndx, &dndy);
        float4 *rec = st.dirRec + kDirRec * (size_t)slot;
        const float4 *src = reinterpret_cast<const float4 *>(&m);
#pragma unroll
        for (int k = 0; k < 6; ++k) rec[k] = src[k];
        rec[6] = mk4(s.ns, 0);
        rec[7] = mk4(s.ss, 0);
        rec[8] = mk4(dpdx, has ? 1.f : 0.f);
        rec[9] = mk4(dpdy, 0);
        rec[10] = mk4(dndx, 0);
        rec[11] = mk4(dndy, 0);
        rec[12] = mk4(rxD, 0);
        rec[13] = mk4(ryD, 0);
    }
}

// The paths of kClassMiss in a scene without infinite lights: nothing to shade — the direct light of the
// previous vertex is folded in and the path is handed to the film kernel (path.cpp:91-101 with no
// infinite light; IISPTdIntegrator's first-hit maps record the miss, iispt_d.cpp:99-115).
__global__ void __launch_bounds__(kBlock) shadeMissKernel(RenderConsts rc, PathState st, Queues q, HemiOutputs hemi) {
    const uint32_t n = q.counters[kClassCounter0 + kClassMiss];
    const uint32_t nLoop = (n + 31u) & ~31u;  // whole warps: queueSlot() is a full-mask vote
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nLoop; i += gridDim.x * blockDim.x) {
        const bool live = i < n;
        uint32_t slot = 0;
        if (live) {
            slot = q.byClass[kClassMiss][i];
            const float4 L4 = st.L[slot];
            const V3f L = f4(L4) + (f4(st.pendS[slot]) + f4(st.pendM[slot]));
            st.L[slot] = mk4(L, L4.w);
            st.pendS[slot] = make_float4(0, 0, 0, 0);
            st.pendM[slot] = make_float4(0, 0, 0, 0);
            if (rc.hemiMode && (__float_as_int(L4.w) & 0xffff) == 0 && hemi.distance) {
                const uint32_t sppLocal = (uint32_t)(rc.sampleEnd - rc.sampleBegin);
                const uint32_t sIdx = st.sample[slot];
                if (sIdx % sppLocal == sppLocal - 1) {  // set_camera_coord overwrites: the pixel's last sample stays
                    const size_t at = (size_t)(sIdx / sppLocal) * rc.hemiSize * rc.hemiSize + st.pixel[slot];
                    hemi.distance[at] = -1.f;
                    hemi.normals[3 * at] = hemi.normals[3 * at + 1] = hemi.normals[3 * at + 2] = 0.f;
                }
            }
        }
        const uint32_t k = queueSlot(&q.counters[4], live);
        if (live) q.done[k] = slot;
    }
}

// ---- shade, Integrator "directlighting" ------------------------------------------------------------
// DirectLightingIntegrator::Li (directlighting.cpp:61-99; the same Li as IILE's
// DirectProgressiveIntegrator, directprogressiveintegrator.cpp:22-62): emission at every hit,
// UniformSampleAllLights / UniformSampleOneLight, and the SpecularReflect / SpecularTransmit recursion
// (integrator.cpp:341-420) while depth + 1 < maxDepth.  The recursion is walked depth-first per
// path, which is the order the reference consumes its sampler in: a vertex whose reflection is
// followed parks itself on the path's stack; when a branch ends the top entry is taken back, its
// surface re-derived, and its transmission sampled with the next two numbers of the stream.
// Sample arrays (integrator.cpp:54-83, directlighting.cpp:44-59): with strategy "all" the integrator asks the
// sampler for two 2D arrays of "samples" points per light and vertex level, maxdepth levels deep, and
// UniformSampleAllLights averages that many EstimateDirect terms per light; when the arrays have run out
// (the specular recursion visits more than maxdepth vertices) it takes one term from Get2D.  The arrays of a
// (pixel, sample) are one more PCG32 stream — sequence = 1 << 62 | sample << 32 | pixel, drawn in the order the
// arrays were requested — which the harness's PixelStreamSampler fills them from; a path finds its place in
// that stream from the number of vertices it has served from arrays (kept in the path's flag word).
// SpecularReflect / SpecularTransmit's ray differentials (integrator.cpp:356-377, :399-427) for the ray leaving
// along wi, from the vertex record textureDirectKernel wrote (PathState::dirRec); out = the path's rayDiff
__device__ __forceinline__ void specularChildDiffs(const float4 *rec, const Surface &s, V3f wi, bool transmit, float bsdfEta,
                                                   float4 *out) {
    const float4 dpdx4 = rec[8];
    if (dpdx4.w == 0.f) {  // ray.hasDifferentials was false
        out[0] = make_float4(0, 0, 0, 0);
        return;
    }
    const V3f dpdx = f4(dpdx4), dpdy = f4(rec[9]), dndx = f4(rec[10]), dndy = f4(rec[11]);
    const V3f wo = s.wo, ns = s.ns;
    const V3f dwodx = -f4(rec[12]) - wo, dwody = -f4(rec[13]) - wo;
    const float dDNdx = dot(dwodx, ns) + dot(wo, dndx);
    const float dDNdy = dot(dwody, ns) + dot(wo, dndy);
    V3f rxD, ryD;
    if (!transmit) {
        rxD = wi - dwodx + 2.f * (dot(wo, ns) * dndx + dDNdx * ns);
        ryD = wi - dwody + 2.f * (dot(wo, ns) * dndy + dDNdy * ns);
    } else {
        float eta = bsdfEta;
        const V3f w = -wo;
        if (dot(wo, ns) < 0) eta = 1.f / eta;
        const float mu = eta * dot(w, ns) - dot(wi, ns);
        const float dmudx = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * dDNdx;
        const float dmudy = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * dDNdy;
        rxD = wi + eta * dwodx - (mu * dndx + dmudx * ns);
        ryD = wi + eta * dwody - (mu * dndy + dmudy * ns);
    }
    out[0] = mk4(s.p + dpdx, 1.f);
    out[1] = mk4(s.p + dpdy, 0);
    out[2] = mk4(rxD, 0);
    out[3] = mk4(ryD, 0);
}

// DIFF: the scene has textures — the vertex's material, shading frame and differentials come from the record
// textureDirectKernel left, and the specular rays carry differentials on
template <bool DIFF>
__global__ void __launch_bounds__(kBlock) shadeDirectKernel(DeviceScene sc, RenderConsts rc, PathState st, Queues q,
                                                            int cur, int strategy) {
    const int nxt = cur ^ 1;
    const uint32_t n = q.counters[cur];
    const uint32_t nLoop = (n + 31u) & ~31u;  // whole warps: queueSlot() is a full-mask vote
    constexpr int nonSpec = BSDF_ALL & ~BSDF_SPECULAR;
    const int nLights = (int)sc.nLights;
    const int nIter = strategy == IILE_DIRECT_ONE ? (nLights > 0 ? 1 : 0) : nLights;
    // floats one vertex takes from the sample arrays (strategy "all"): 4 per sample of every light
    uint32_t arrayFloatsPerVertex = 0;
    if (strategy == IILE_DIRECT_ALL)
        for (int j = 0; j < nLights; ++j) arrayFloatsPerVertex += 4u * (uint32_t)IILE_LIGHT_NSAMPLES(sc.lights[j]);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nLoop; i += gridDim.x * blockDim.x) {
        const bool live = i < n;
        uint32_t slot = 0;
        bool wantNext = false, wantDone = false, shadeIt = false, needPop = false;
        float4 o4 = make_float4(0, 0, 0, 0), d4 = o4;
        V3f o, d, beta, L;
        int depth = 0, nStack = 0, leafSlot = -1;
        int arrayVertices = 0;  // vertices of this path that took their light samples from the arrays
        int spawned = 0;        // DIFF: the ray is no longer the camera's (its differentials, if any, are in rayDiff)
        SampleStream g;
        Surface s;
        Bsdf bsdf;
        if (live) {
            slot = q.active[cur][i];
            o4 = st.rayO[slot];
            d4 = st.rayD[slot];
            const float4 beta4 = st.beta[slot], L4 = st.L[slot];
            const int2 hit = st.hit[slot];
            o = f4(o4), d = f4(d4), beta = f4(beta4), L = f4(L4);
            const int flags = __float_as_int(L4.w);
            depth = flags & 0xffff;
            nStack = (flags >> 20) & 0xf;
            arrayVertices = (flags >> 24) & 0xf;
            spawned = (flags >> 28) & 1;
            leafSlot = hit.x;
            g.unpack(st.rng[slot], st.sample[slot], st.pixel[slot], rc.halton);
            float tHit = 0;
            const bool found = hit.x >= 0 && surfaceAt(sc, hit.x, o, d, &s, &tHit);
            if (!found) {  // directlighting.cpp:67-70: every light's Le(ray)
                if (sc.nInfinite > 0)
                    for (uint32_t k = 0; k < sc.nLights; ++k)
                        if (sc.lights[k].type == IILE_LIGHT_INFINITE) L = L + beta * envLe(sc.lights[k], d);
                needPop = true;
            } else if (s.material < 0) {  // no BSDF: Li(isect.SpawnRay(ray.d), ..., depth), directlighting.cpp:74-75
                st.rayO[slot] = mk4(offsetRayOrigin(s.p, s.pError, s.n, d), o4.w);
                if constexpr (DIFF) {  // a plain Ray: no differentials from here on
                    st.rayDiff[4 * (size_t)slot] = make_float4(0, 0, 0, 0);
                    spawned = 1;
                }
                wantNext = true;
            } else {
                L = L + beta * emitted(sc, s.light, s.n, -d);  // isect.Le(wo)
                const IileMaterial *mp = &sc.materials[s.material];
                if constexpr (DIFF) {
                    const float4 *rec = st.dirRec + kDirRec * (size_t)slot;
                    mp = reinterpret_cast<const IileMaterial *>(rec);
                    s.ns = f4(rec[6]);
                    s.ss = f4(rec[7]);
                    s.ts = cross(s.ns, s.ss);
                }
                makeBsdf(*mp, s, &bsdf, false);
                shadeIt = true;
            }
        }
        // ---- direct lighting: the same trip count for every lane of the warp -------------------------
        // this vertex's light samples come from the arrays while the path has arrays left (maxdepth levels)
        const bool fromArrays = strategy == IILE_DIRECT_ALL && arrayVertices < rc.maxDepth;
        uint32_t arrayAt = (uint32_t)arrayVertices * arrayFloatsPerVertex;  // this vertex's place in the array stream
        // GlobalSampler (halton / sobol) arrays (sampler.cpp:128-162): 2D array i lives in dimensions 5 + 2i and
        // 5 + 2i + 1; element k of pixel sample s is drawn at the pixel's sample number s * n + k
        const bool globalArrays = rc.halton.tables || rc.halton.sobol.matrices;
        int apx = 0, apy = 0;
        Pcg32 arrayBase;
        if (fromArrays && shadeIt) {
            if (globalArrays) {
                const int sbw = rc.sampleBounds[2] - rc.sampleBounds[0];
                apx = (int)(st.pixel[slot] % (uint32_t)sbw);  // relative to the sample bounds' corner
                apy = (int)(st.pixel[slot] / (uint32_t)sbw);
            } else
                arrayBase.setSequence(((uint64_t)1 << 62) | ((uint64_t)st.sample[slot] << 32) | st.pixel[slot]);
        }
        for (int j = 0; j < nIter; ++j) {
          // RoundCount is the identity for every sampler in scope (this tree's SobolSampler does not override it)
          const int nSamples = strategy == IILE_DIRECT_ALL ? IILE_LIGHT_NSAMPLES(sc.lights[j]) : 1;
          const int nTerms = fromArrays ? nSamples : 1;  // integrator.cpp:66-72: without arrays, one term
          Pcg32 arrL = arrayBase, arrS = arrayBase;
          if (fromArrays && shadeIt) {
              arrL.advance((int64_t)arrayAt);                       // uLightArray of light j
              arrS.advance((int64_t)arrayAt + 2 * (int64_t)nSamples);  // uScatteringArray of light j
          }
          arrayAt += 4u * (uint32_t)nSamples;
          for (int term = 0; term < nSamples; ++term) {  // warp-uniform trip count: queueSlot() is a full-mask vote
            bool wantShadow = false, wantMis = false;
            ShadowItem sh;
            MisItem mi;
            float4 gated = make_float4(0, 0, 0, 0);
            if (shadeIt && term < nTerms) {
                int lightNum = j;
                float pickPdf = 1.f;
                if (strategy == IILE_DIRECT_ONE) {  // UniformSampleOneLight without a distribution, integrator.cpp:93-96
                    lightNum = min((int)(g.uniformFloat() * nLights), nLights - 1);
                    pickPdf = 1.f / nLights;
                }
                if (fromArrays) pickPdf = (float)nTerms;  // Ld / nSamples (integrator.cpp:80): folded into every term
                float ul0, ul1, uaS0 = 0, uaS1 = 0;
                if (fromArrays && globalArrays) {
                    const int a = 2 * (arrayVertices * nLights + j);  // the uLight array; the uScattering array is a + 1
                    const uint32_t sampleNum = st.sample[slot] * (uint32_t)nSamples + (uint32_t)term;
                    if (rc.halton.tables) {
                        const uint64_t idx = haltonIndexForSample(rc.halton, apx + rc.sampleBounds[0], apy + rc.sampleBounds[1], sampleNum);
                        ul0 = haltonScrambled(rc.halton.tables, idx, 5 + 2 * a);
                        ul1 = haltonScrambled(rc.halton.tables, idx, 5 + 2 * a + 1);
                        uaS0 = haltonScrambled(rc.halton.tables, idx, 5 + 2 * a + 2);
                        uaS1 = haltonScrambled(rc.halton.tables, idx, 5 + 2 * a + 3);
                    } else {
                        const uint64_t idx = sobolIntervalToIndex(rc.halton.sobol, sampleNum, (uint32_t)apx, (uint32_t)apy);
                        ul0 = sobolSample(rc.halton.sobol.matrices, idx, 5 + 2 * a);
                        ul1 = sobolSample(rc.halton.sobol.matrices, idx, 5 + 2 * a + 1);
                        uaS0 = sobolSample(rc.halton.sobol.matrices, idx, 5 + 2 * a + 2);
                        uaS1 = sobolSample(rc.halton.sobol.matrices, idx, 5 + 2 * a + 3);
                    }
                } else {
                    ul0 = fromArrays ? arrL.uniformFloat() : g.uniformFloat();
                    ul1 = fromArrays ? arrL.uniformFloat() : g.uniformFloat();
                }
                const IileLight &light = sc.lights[lightNum];
                LightSample ls;
                sampleLight(sc, light, s.p, s.pError, s.n, ul0, ul1, &ls);
                if (strategy == 2) {  // WhittedIntegrator's loop, whitted.cpp:70-81: no MIS, f over every lobe
                    if (!isBlack(ls.Li) && ls.pdf != 0) {
                        const V3f f = bsdfF(bsdf, s.wo, ls.wi, BSDF_ALL);
                        if (!isBlack(f)) {
                            const V3f so = ls.pOrigin, sd = ls.pTarget - so;
                            sh.o = mk4(so, __int_as_float((int)slot));
                            sh.d = mk4(sd, 1 - IILE_SHADOW_EPS);
                            gated = mk4(beta * sdiv(f * ls.Li * absDot(ls.wi, s.ns), ls.pdf), 0);
                            wantShadow = true;
                        }
                    }
                }
                const float us0 = strategy == 2 ? 0.f : (fromArrays ? (globalArrays ? uaS0 : arrS.uniformFloat()) : g.uniformFloat());
                const float us1 = strategy == 2 ? 0.f : (fromArrays ? (globalArrays ? uaS1 : arrS.uniformFloat()) : g.uniformFloat());
                const bool delta = strategy == 2 || (light.type != IILE_LIGHT_DIFFUSE_AREA && light.type != IILE_LIGHT_INFINITE);
                if (strategy != 2 && ls.pdf > 0 && !isBlack(ls.Li)) {
                    V3f f = bsdfF(bsdf, s.wo, ls.wi, nonSpec) * absDot(ls.wi, s.ns);
                    const float scatteringPdf = bsdfPdf(bsdf, s.wo, ls.wi, nonSpec);
                    if (!isBlack(f)) {
                        V3f Ld = delta ? sdiv(f * ls.Li, ls.pdf)
                                       : sdiv(f * ls.Li * powerHeuristic(ls.pdf, scatteringPdf), ls.pdf);
                        if (strategy == IILE_DIRECT_ONE || fromArrays) Ld = sdiv(Ld, pickPdf);
                        const V3f so = ls.pOrigin, sd = ls.pTarget - so;
                        sh.o = mk4(so, __int_as_float((int)slot));
                        sh.d = mk4(sd, 1 - IILE_SHADOW_EPS);
                        gated = mk4(beta * Ld, 0);
                        wantShadow = true;
                    }
                }
                if (!delta) {
                    V3f wi;
                    float scatteringPdf;
                    int sampledType;
                    V3f f = bsdfSample(bsdf, s.wo, &wi, us0, us1, &scatteringPdf, nonSpec, &sampledType);
                    if (scatteringPdf > 0) f = f * absDot(wi, s.ns);
                    if (!isBlack(f) && scatteringPdf > 0) {
                        const V3f ro = offsetRayOrigin(s.p, s.pError, s.n, wi);
                        const float lightPdf = light.type == IILE_LIGHT_INFINITE
                                                   ? envPdf(light, wi)
                                                   : areaLightPdf(sc, light, s.p, s.pError, s.n, ro, wi);
                        if (lightPdf != 0) {
                            const float weight = powerHeuristic(scatteringPdf, lightPdf);
                            V3f c = sdiv(f * weight, scatteringPdf);
                            if (strategy == IILE_DIRECT_ONE || fromArrays) c = sdiv(c, pickPdf);
                            mi.o = mk4(ro, __int_as_float((int)slot));
                            mi.d = mk4(wi, __int_as_float(lightNum));
                            mi.contrib = mk4(beta * c, __int_as_float(-1));
                            wantMis = true;
                        }
                    }
                }
            }
            uint32_t k = queueSlot(&q.counters[2], wantShadow);
            if (wantShadow) {
                q.shadow[k] = sh;
                q.shadowContrib[k] = gated;
            }
            k = queueSlot(&q.counters[3], wantMis);
            if (wantMis) q.mis[k] = mi;
          }
        }
        if (shadeIt && fromArrays && nLights > 0) ++arrayVertices;
        // ---- SpecularReflect, then SpecularTransmit (now, or when the reflected branch has ended) ----
        if (live) {
            constexpr int kRefl = BSDF_REFLECTION | BSDF_SPECULAR, kTrans = BSDF_TRANSMISSION | BSDF_SPECULAR;
            if (shadeIt) {
                if (depth + 1 < rc.maxDepth) {
                    V3f wi;
                    float pdf;
                    int sampled;
                    float u0 = g.uniformFloat(), u1 = g.uniformFloat();
                    V3f f = bsdfSample(bsdf, s.wo, &wi, u0, u1, &pdf, kRefl, &sampled);
                    if (pdf > 0.f && !isBlack(f) && absDot(wi, s.ns) != 0.f) {
                        if (nStack < kDirectStack) {  // its transmission waits for the reflected branch
                            const size_t at = (size_t)slot * kDirectStack + nStack;
                            st.stackO[at] = mk4(o, __int_as_float(depth));
                            st.stackD[at] = mk4(d, __int_as_float(leafSlot));
                            st.stackBeta[at] = mk4(beta, 0);
                            if constexpr (DIFF) {
                                const float4 *rec = st.dirRec + kDirRec * (size_t)slot;
                                float4 *keep = st.stackRec + kDirRec * at;
#pragma unroll
                                for (int r = 0; r < kDirRec; ++r) keep[r] = rec[r];
                            }
                            ++nStack;
                        }
                        beta = beta * sdiv(f * absDot(wi, s.ns), pdf);
                        st.rayO[slot] = mk4(offsetRayOrigin(s.p, s.pError, s.n, wi), o4.w);
                        st.rayD[slot] = mk4(wi, d4.w);
                        if constexpr (DIFF) {
                            specularChildDiffs(st.dirRec + kDirRec * (size_t)slot, s, wi, false, bsdf.eta, st.rayDiff + 4 * (size_t)slot);
                            spawned = 1;
                        }
                        ++depth;
                        wantNext = true;
                    } else {
                        u0 = g.uniformFloat(), u1 = g.uniformFloat();
                        f = bsdfSample(bsdf, s.wo, &wi, u0, u1, &pdf, kTrans, &sampled);
                        if (pdf > 0.f && !isBlack(f) && absDot(wi, s.ns) != 0.f) {
                            beta = beta * sdiv(f * absDot(wi, s.ns), pdf);
                            st.rayO[slot] = mk4(offsetRayOrigin(s.p, s.pError, s.n, wi), o4.w);
                            st.rayD[slot] = mk4(wi, d4.w);
                            if constexpr (DIFF) {
                                specularChildDiffs(st.dirRec + kDirRec * (size_t)slot, s, wi, true, bsdf.eta, st.rayDiff + 4 * (size_t)slot);
                                spawned = 1;
                            }
                            ++depth;
                            wantNext = true;
                        } else
                            needPop = true;
                    }
                } else
                    needPop = true;
            }
            while (needPop) {
                if (nStack == 0) {
                    wantDone = true;
                    break;
                }
                --nStack;
                const size_t at = (size_t)slot * kDirectStack + nStack;
                const float4 eo = st.stackO[at], ed = st.stackD[at], eb = st.stackBeta[at];
                float tHit = 0;
                if (!surfaceAt(sc, __float_as_int(ed.w), f4(eo), f4(ed), &s, &tHit)) continue;
                const IileMaterial *mp = &sc.materials[s.material];
                if constexpr (DIFF) {
                    const float4 *rec = st.stackRec + kDirRec * at;
                    mp = reinterpret_cast<const IileMaterial *>(rec);
                    s.ns = f4(rec[6]);
                    s.ss = f4(rec[7]);
                    s.ts = cross(s.ns, s.ss);
                }
                makeBsdf(*mp, s, &bsdf, false);
                V3f wi;
                float pdf;
                int sampled;
                const float u0 = g.uniformFloat(), u1 = g.uniformFloat();
                V3f f = bsdfSample(bsdf, s.wo, &wi, u0, u1, &pdf, kTrans, &sampled);
                if (pdf > 0.f && !isBlack(f) && absDot(wi, s.ns) != 0.f) {
                    beta = f4(eb) * sdiv(f * absDot(wi, s.ns), pdf);
                    st.rayO[slot] = mk4(offsetRayOrigin(s.p, s.pError, s.n, wi), o4.w);
                    st.rayD[slot] = mk4(wi, d4.w);
                    if constexpr (DIFF) {
                        specularChildDiffs(st.stackRec + kDirRec * at, s, wi, true, bsdf.eta, st.rayDiff + 4 * (size_t)slot);
                        spawned = 1;
                    }
                    depth = __float_as_int(eo.w) + 1;
                    wantNext = true;
                    needPop = false;
                }
            }
            st.beta[slot] = mk4(beta, 1);
            st.L[slot] = mk4(L, __int_as_float((depth & 0xffff) | (nStack << 20) | (arrayVertices << 24) | (spawned << 28)));
            st.rng[slot] = g.pack();
        }
        uint32_t k = queueSlot(&q.counters[nxt], wantNext);
        if (wantNext) q.active[nxt][k] = slot;
        k = queueSlot(&q.counters[4], wantDone);
        if (wantDone) q.done[k] = slot;
    }
}

// ---- ray binning: counting sort of the active queue by (origin cell, direction octant) -------------
__device__ __forceinline__ uint32_t spread5(uint32_t x) {  // 5 bits -> every third bit
    x &= 0x1fu;
    x = (x | (x << 8)) & 0x100fu;
    x = (x | (x << 4)) & 0x10c3u;
    x = (x | (x << 2)) & 0x1249u;
    return x;
}
__global__ void __launch_bounds__(kBlock) sortKeyKernel(DeviceScene sc, PathState st, Queues q, int cur,
                                                        uint32_t *__restrict__ keys, uint32_t *__restrict__ hist) {
    const uint32_t n = q.counters[cur];
    const float sx = 32.f / fmaxf(sc.worldMax[0] - sc.worldMin[0], 1e-20f);
    const float sy = 32.f / fmaxf(sc.worldMax[1] - sc.worldMin[1], 1e-20f);
    const float sz = 32.f / fmaxf(sc.worldMax[2] - sc.worldMin[2], 1e-20f);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = q.active[cur][i];
        const float4 o = st.rayO[slot], d = st.rayD[slot];
        const uint32_t cx = (uint32_t)fminf(fmaxf((o.x - sc.worldMin[0]) * sx, 0.f), 31.f);
        const uint32_t cy = (uint32_t)fminf(fmaxf((o.y - sc.worldMin[1]) * sy, 0.f), 31.f);
        const uint32_t cz = (uint32_t)fminf(fmaxf((o.z - sc.worldMin[2]) * sz, 0.f), 31.f);
        const uint32_t octant = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
        const uint32_t key = ((spread5(cx) | (spread5(cy) << 1) | (spread5(cz) << 2)) << 3) | octant;
        keys[i] = key;
        atomicAdd(&hist[key], 1u);
    }
}
// counts -> exclusive offsets, one block
__global__ void __launch_bounds__(1024) sortScanKernel(uint32_t *hist) {
    __shared__ uint32_t partial[1024];
    constexpr int per = kSortBins / 1024;
    const int t = threadIdx.x;
    uint32_t sum = 0;
    for (int k = 0; k < per; ++k) sum += hist[t * per + k];
    partial[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t v = t >= off ? partial[t - off] : 0u;
        __syncthreads();
        partial[t] += v;
        __syncthreads();
    }
    uint32_t run = partial[t] - sum;
    for (int k = 0; k < per; ++k) {
        const uint32_t c = hist[t * per + k];
        hist[t * per + k] = run;
        run += c;
    }
}
__global__ void __launch_bounds__(kBlock) sortScatterKernel(Queues q, int cur, const uint32_t *__restrict__ keys,
                                                            uint32_t *__restrict__ cursor, uint32_t *__restrict__ sorted) {
    const uint32_t n = q.counters[cur];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        sorted[atomicAdd(&cursor[keys[i]], 1u)] = q.active[cur][i];
}
__global__ void __launch_bounds__(kBlock) sortCopyKernel(Queues q, int cur, const uint32_t *__restrict__ sorted) {
    const uint32_t n = q.counters[cur];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        q.active[cur][i] = sorted[i];
}

// ---- cast: next rays + MIS rays + shadow rays in one launch ---------------------------------------
// Item space: [0, nNext) closest hits of active[cur]; [nNext, nNext+nMis) MIS rays; the rest shadow
// rays.  Long queries first, so the launch's tail is made of the short any-hit rays.
// References to the kernel's __grid_constant__ parameters: queue pointers are read from the
// constant bank where they are used instead of living in registers across the traversal loop.
// A MIS ray that reaches the light it was sampled for only leaves the leaf slot in its item
// (contrib.w); misResolveKernel evaluates Le afterwards — a call to the surface code from inside
// the traversal loop costs the loop ~20 registers (spills at 7 blocks/SM).
struct WaveRaySrc {
    const DeviceScene &scr;
    const PathState &str;
    const Queues &qr;
    int cur;
    uint32_t nNext, nMis;
    __device__ __forceinline__ bool load(uint32_t i, RayCtx &r, bool &any) const {
        float4 o, d;
        float tMax = INFINITY;
        any = false;
        if (i < nNext) {
            const uint32_t slot = qr.active[cur][i];
            o = LD_STATE(str.rayO + slot);
            d = LD_STATE(str.rayD + slot);
        } else if (i < nNext + nMis) {
            o = LD_STATE(&qr.mis[i - nNext].o);
            d = LD_STATE(&qr.mis[i - nNext].d);
        } else {
            const ShadowItem *it = qr.shadow + (i - nNext - nMis);
            o = LD_STATE(&it->o);
            d = LD_STATE(&it->d);
            tMax = d.w;
            any = true;
        }
        initRay(r, o.x, o.y, o.z, d.x, d.y, d.z, tMax);
        return true;
    }
    __device__ __forceinline__ void prefetch(uint32_t, uint32_t, unsigned) const {}
    __device__ __forceinline__ void store(uint32_t i, const RayCtx &, const HitRec &h, bool found) const {
        if (i < nNext) {
            ST_STATE(str.hit + qr.active[cur][i], make_int2(h.slot, __float_as_int(h.slot >= 0 ? h.t : 0.f)));
        } else if (i < nNext + nMis) {
            if (h.slot < 0) {
                // escaped: Li = light.Le(ray), non-zero only for an infinite light (integrator.cpp:207-208)
                if (scr.nInfinite > 0 && scr.lights[__float_as_int(qr.mis[i - nNext].d.w)].type == IILE_LIGHT_INFINITE)
                    qr.mis[i - nNext].contrib.w = __int_as_float(-2);
                return;
            }
            const MisItem *it = qr.mis + (i - nNext);
            const int primIdx = __float_as_int(__ldg(scr.leafPrims + 3 * h.slot).w);
            if (scr.prims[primIdx].light != __float_as_int(it->d.w)) return;
            qr.mis[i - nNext].contrib.w = __int_as_float(h.slot);
        } else {
            const uint32_t k = i - nNext - nMis;
            if (qr.directMode) {  // unoccluded: the gated term joins the path's radiance
                if (found) return;
                const uint32_t slot = (uint32_t)__float_as_int(qr.shadow[k].o.w);
                const float4 c = qr.shadowContrib[k];
                float *dst = reinterpret_cast<float *>(str.L + slot);
                atomicAdd(dst, c.x);
                atomicAdd(dst + 1, c.y);
                atomicAdd(dst + 2, c.z);
                return;
            }
            if (!found) return;
            const uint32_t slot = (uint32_t)__float_as_int(qr.shadow[k].o.w);
            str.pendS[slot] = make_float4(0, 0, 0, 0);
        }
    }
};

// cur < 0: only the MIS and shadow queues (after the last shade of a batch)
__global__ void __launch_bounds__(kCastBlock, kCastMinBlocks) castKernel(const __grid_constant__ DeviceScene sc,
                                                             const __grid_constant__ PathState st,
                                                             const __grid_constant__ Queues q, int cur) {
    const uint32_t nNext = cur >= 0 ? q.counters[cur] : 0u, nMis = q.counters[3], nShadow = q.counters[2];
    WaveRaySrc src{sc, st, q, cur >= 0 ? cur : 0, nNext, nMis};
    castPersistent<kCastMixed>(sc, src, nNext + nMis + nShadow, sc.castCounters);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&q.stats[0], (unsigned long long)nNext + nMis);
        atomicAdd(&q.stats[1], (unsigned long long)nShadow);
    }
    releaseWorkCounter(sc.castCounters);
}

// Measurement aid (iile_render_dump_rays_dev): copies every `stride`-th ray the next cast launch will
// trace — item space as in WaveRaySrc — into a caller buffer, with its query kind.
__global__ void __launch_bounds__(kBlock) dumpRaysKernel(PathState st, Queues q, int cur, RayDump dump) {
    const uint32_t nNext = cur >= 0 ? q.counters[cur] : 0u, nMis = q.counters[3], nShadow = q.counters[2];
    const uint32_t n = nNext + nMis + nShadow;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i % dump.stride != dump.phase % dump.stride) continue;
        float4 o, d;
        float tMax = INFINITY;
        uint8_t kind = 0;
        if (i < nNext) {
            const uint32_t slot = q.active[cur >= 0 ? cur : 0][i];
            o = st.rayO[slot];
            d = st.rayD[slot];
            kind = (__float_as_int(st.L[slot].w) & 0xffff) == 0 ? 0 : 1;  // camera ray | bounce ray
        } else if (i < nNext + nMis) {
            o = q.mis[i - nNext].o;
            d = q.mis[i - nNext].d;
            kind = 2;
        } else {
            o = q.shadow[i - nNext - nMis].o;
            d = q.shadow[i - nNext - nMis].d;
            tMax = d.w;
            kind = 3;
        }
        const unsigned long long at = atomicAdd(dump.count, 1ull);
        if (at >= dump.cap) continue;
        float4 *dst = reinterpret_cast<float4 *>(dump.rays + at);
        dst[0] = make_float4(o.x, o.y, o.z, tMax);
        dst[1] = make_float4(d.x, d.y, d.z, 0.f);
        dump.kinds[at] = kind;
    }
}

// path: one MIS term per path and vertex, parked in pendM; directlighting: one per light, summed into L
__device__ __forceinline__ void creditMis(const PathState &st, const Queues &q, uint32_t slot, float4 c) {
    if (q.directMode) {
        float *dst = reinterpret_cast<float *>(st.L + slot);
        atomicAdd(dst, c.x);
        atomicAdd(dst + 1, c.y);
        atomicAdd(dst + 2, c.z);
    } else
        st.pendM[slot] = c;
}

// MIS rays that reached their light: Le through the power-heuristic weight (integrator.cpp:196-211).
__global__ void __launch_bounds__(kBlock) misResolveKernel(DeviceScene sc, PathState st, Queues q) {
    const uint32_t n = q.counters[3];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int leafSlot = __float_as_int(q.mis[i].contrib.w);
        if (leafSlot < 0 && leafSlot != -2) continue;
        const MisItem it = q.mis[i];
        const int light = __float_as_int(it.d.w);
        const V3f o = f4(it.o), d = f4(it.d);
        if (leafSlot == -2) {  // the BSDF-sampled ray escaped towards its infinite light
            const V3f Le = envLe(sc.lights[light], d);
            if (isBlack(Le)) continue;
            const uint32_t slot = (uint32_t)__float_as_int(it.o.w);
            creditMis(st, q, slot, make_float4(it.contrib.x * Le.x, it.contrib.y * Le.y, it.contrib.z * Le.z, 0));
            continue;
        }
        Surface s;
        float t;
        if (!surfaceAt(sc, leafSlot, o, d, &s, &t)) continue;
        const V3f Le = emitted(sc, light, s.n, -d);
        if (isBlack(Le)) continue;
        const uint32_t slot = (uint32_t)__float_as_int(it.o.w);
        creditMis(st, q, slot, make_float4(it.contrib.x * Le.x, it.contrib.y * Le.y, it.contrib.z * Le.z, 0));
    }
}

// ---- film ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) filmKernel(RenderConsts rc, PathState st, Queues q, float4 *film,
                                                     float *hemiIntensity) {
    const uint32_t n = q.counters[4];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = q.done[i];
        const float4 L4 = st.L[slot];
        V3f L = f4(L4) + (f4(st.pendS[slot]) + f4(st.pendM[slot]));
        // integrator.cpp:293-314
        const float y = luminance(L);
        if (isnan(L.x) || isnan(L.y) || isnan(L.z)) L = V3f(0, 0, 0);
        else if (y < -1e-5f) L = V3f(0, 0, 0);
        else if (isinf(y)) L = V3f(0, 0, 0);
        if (rc.filmMonitor) {
            // IisptFilmMonitor::add_n_samples: RGB and ray weight summed on the pixel of the loop, not on pFilm's
            const uint32_t sbw = (uint32_t)(rc.sampleBounds[2] - rc.sampleBounds[0]);
            const int mx = rc.sampleBounds[0] + (int)(st.pixel[slot] % sbw), my = rc.sampleBounds[1] + (int)(st.pixel[slot] / sbw);
            if (mx >= rc.cropped[0] && mx < rc.cropped[2] && my >= rc.cropped[1] && my < rc.cropped[3])
                atomicAdd(film + (size_t)(my - rc.cropped[1]) * (rc.cropped[2] - rc.cropped[0]) + (mx - rc.cropped[0]),
                          make_float4(L.x, L.y, L.z, 1.f));
            continue;
        }
        // FilmTile::AddSample, film.h:153-193
        const float ly = luminance(L);
        if (ly > rc.maxSampleLuminance) L = L * (rc.maxSampleLuminance / ly);
        const float pfx = st.rayO[slot].w, pfy = st.rayD[slot].w;
        const float dx = pfx - 0.5f, dy = pfy - 0.5f;
        int x0 = (int)ceilf(dx - rc.filterRadius[0]), x1 = (int)floorf(dx + rc.filterRadius[0]) + 1;
        int y0 = (int)ceilf(dy - rc.filterRadius[1]), y1 = (int)floorf(dy + rc.filterRadius[1]) + 1;
        x0 = max(x0, rc.cropped[0]);
        x1 = min(x1, rc.cropped[2]);
        y0 = max(y0, rc.cropped[1]);
        y1 = min(y1, rc.cropped[3]);
        const int fw = rc.cropped[2] - rc.cropped[0];
        float4 *base = film;
        if (rc.hemiMode) {
            const uint32_t sppLocal = (uint32_t)(rc.sampleEnd - rc.sampleBegin);
            base = film + (size_t)(st.sample[slot] / sppLocal) * rc.hemiSize * rc.hemiSize;
        }
        for (int yy = y0; yy < y1; ++yy) {
            float fy = fabsf((yy - dy) * rc.invFilterRadius[1] * 16);
            int iy = min((int)floorf(fy), 15);
            for (int xx = x0; xx < x1; ++xx) {
                float fx = fabsf((xx - dx) * rc.invFilterRadius[0] * 16);
                int ix = min((int)floorf(fx), 15);
                const float wgt = rc.filterTable[iy * 16 + ix];
                V3f c = L * 1.f * wgt;  // L * sampleWeight * filterWeight
                float4 *px = base + (size_t)(yy - rc.cropped[1]) * fw + (xx - rc.cropped[0]);
                atomicAdd(px, make_float4(c.x, c.y, c.z, wgt));
            }
        }
    }
}

__global__ void resetCountersKernel(Queues q, int next) {
    q.counters[next] = 0;
    q.counters[2] = 0;
    q.counters[3] = 0;
    q.counters[4] = 0;
    for (int c = 0; c < kNumClasses; ++c) q.counters[kClassCounter0 + c] = 0;
}

// ---- spatial light grid ---------------------------------------------------------------------------
// One thread per (voxel, light) accumulates the 128 Halton-point estimates in order
// (lightdistrib.cpp:254-279); one thread per voxel then builds the Distribution1D
// (lightdistrib.cpp:281-298, sampling.h:57-69).
__global__ void lightGridAccumulate(DeviceScene sc, const float *__restrict__ halton, float *__restrict__ contrib) {
    const size_t nVox = (size_t)sc.voxels[0] * sc.voxels[1] * sc.voxels[2];
    const size_t total = nVox * sc.nLights;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t voxel = i / sc.nLights;
        const int j = (int)(i % sc.nLights);
        int vi[3];
        vi[0] = (int)(voxel % sc.voxels[0]);
        vi[1] = (int)((voxel / sc.voxels[0]) % sc.voxels[1]);
        vi[2] = (int)(voxel / ((size_t)sc.voxels[0] * sc.voxels[1]));
        V3f lo, hi;
        {
            float l[3], h[3];
            for (int a = 0; a < 3; ++a) {
                float t0 = (float)vi[a] / (float)sc.voxels[a], t1 = (float)(vi[a] + 1) / (float)sc.voxels[a];
                float p0 = (1 - t0) * sc.worldMin[a] + t0 * sc.worldMax[a];
                float p1 = (1 - t1) * sc.worldMin[a] + t1 * sc.worldMax[a];
                l[a] = fminf(p0, p1);
                h[a] = fmaxf(p0, p1);
            }
            lo = V3f(l[0], l[1], l[2]);
            hi = V3f(h[0], h[1], h[2]);
        }
        float acc = 0;
        for (int k = 0; k < 128; ++k) {
            const float *hp = halton + 5 * k;
            V3f po((1 - hp[0]) * lo.x + hp[0] * hi.x, (1 - hp[1]) * lo.y + hp[1] * hi.y,
                   (1 - hp[2]) * lo.z + hp[2] * hi.z);
            LightSample ls;
            sampleLight(sc, sc.lights[j], po, V3f(0, 0, 0), V3f(0, 0, 0), hp[3], hp[4], &ls);
            if (ls.pdf > 0) acc += luminance(ls.Li) / ls.pdf;
        }
        contrib[i] = acc;
    }
}

__global__ void lightGridFinalize(DeviceScene sc, const float *__restrict__ contrib, float *__restrict__ table) {
    const size_t nVox = (size_t)sc.voxels[0] * sc.voxels[1] * sc.voxels[2];
    const int n = (int)sc.nLights;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nVox; v += (size_t)gridDim.x * blockDim.x) {
        const float *c = contrib + v * n;
        float *cdf = table + v * (2 * (size_t)n + 2), *func = cdf + n + 1;
        float sum = 0;
        for (int i = 0; i < n; ++i) sum = sum + c[i];
        float avg = sum / (128 * n);
        float minC = (avg > 0) ? .001f * avg : 1;
        cdf[0] = 0;
        for (int i = 0; i < n; ++i) func[i] = fmaxf(c[i], minC);
        for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / n;
        float funcInt = cdf[n];
        if (funcInt == 0) {
            for (int i = 1; i < n + 1; ++i) cdf[i] = (float)i / (float)n;
        } else {
            for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
        }
        cdf[2 * n + 1] = funcInt;
    }
}

}  // namespace

#define LAUNCH(kernel, l, ...)                                    \
    kernel<<<(l).blocks, kBlock, 0, (l).stream>>>(__VA_ARGS__);   \
    return cudaGetLastError();

cudaError_t wfGenerate(const WavefrontLaunch &l, const RenderConsts &rc, PathState st, Queues q,
                       uint64_t firstWork, uint32_t nWork, HemiOutputs hemi) {
    LAUNCH(generateKernel, l, rc, st, q, firstWork, nWork, hemi)
}
// traversal kernels: persistent grid sized by occupancy
template <typename K>
static int castBlocks(K kernel) {
    int dev = 0, sms = 148, perSm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kCastBlock, 0);
    return sms * (perSm > 0 ? perSm : 1);
}
#define LAUNCH_CAST(kernel, l, ...)                                              \
    static const int blocks = castBlocks(kernel);                                \
    kernel<<<blocks, kCastBlock, 0, (l).stream>>>(__VA_ARGS__);          \
    return cudaGetLastError();

cudaError_t wfCast(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q, int cur) {
    static const int blocks = castBlocks(castKernel);
    castKernel<<<blocks, kCastBlock, 0, l.stream>>>(sc, st, q, cur);
    return cudaGetLastError();
}
cudaError_t wfDumpRays(const WavefrontLaunch &l, PathState st, Queues q, int cur, RayDump dump) {
    LAUNCH(dumpRaysKernel, l, st, q, cur, dump)
}
cudaError_t wfMisResolve(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q) {
    LAUNCH(misResolveKernel, l, sc, st, q)
}
cudaError_t wfClassify(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q, int cur) {
    classifyKernel<<<l.blocks / 4, kClassifySeg, 0, l.stream>>>(sc, st, q, cur);
    return cudaGetLastError();
}
// One generic launch over the classes of `genericMask`, plus one launch for each of kClassMiss /
// kClassMatte / kClassPlastic named in `fastMask` (the caller names them only for scenes whose lights
// the per-material kernels sample in line: point lights and diffuse area lights on triangles).
cudaError_t wfShade(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st,
                    Queues q, int cur, HemiOutputs hemi, int genericMask, int fastMask, bool simpleLights,
                    uint64_t *launches) {
    // one large block per SM slot (see the phase barriers): grid = SM count x resident blocks
    auto perSm = [](auto kernel, int threads) {
        int v = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, 0);
        return v > 0 ? v : 1;
    };
    const int sms = l.blocks / 8;
    if (genericMask) {
        static const int per = perSm(shadeKernelT<-1, false>, kShadeBlock);
        shadeKernelT<-1, false><<<sms * per, kShadeBlock, 0, l.stream>>>(sc, rc, st, q, cur, hemi, genericMask);
        ++*launches;
    }
#define FAST(MAT, CLS)                                                                                          \
    if ((fastMask >> (CLS)) & 1) {                                                                              \
        static const int per = perSm(shadeKernelT<MAT, true>, kFastBlock);                                      \
        shadeKernelT<MAT, true><<<sms * per, kFastBlock, 0, l.stream>>>(sc, rc, st, q, cur, hemi, 1 << (CLS));  \
        ++*launches;                                                                                            \
    }
    FAST(IILE_MAT_MATTE, kClassMatte)
    FAST(IILE_MAT_PLASTIC, kClassPlastic)
    if ((fastMask >> kClassMiss) & 1) {
        shadeMissKernel<<<l.blocks, kBlock, 0, l.stream>>>(rc, st, q, hemi);
        ++*launches;
    }
#undef FAST
    return cudaGetLastError();
}
cudaError_t wfTextures(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st, Queues q,
                       HemiOutputs hemi) {
    textureKernel<<<l.blocks, kBlock, 0, l.stream>>>(sc, rc, st, q, hemi);
    return cudaGetLastError();
}

cudaError_t wfSortActive(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q, int cur,
                         uint32_t *keys, uint32_t *hist, uint32_t *sorted) {
    cudaError_t e = cudaMemsetAsync(hist, 0, kSortBins * sizeof(uint32_t), l.stream);
    if (e != cudaSuccess) return e;
    sortKeyKernel<<<l.blocks, kBlock, 0, l.stream>>>(sc, st, q, cur, keys, hist);
    sortScanKernel<<<1, 1024, 0, l.stream>>>(hist);
    sortScatterKernel<<<l.blocks, kBlock, 0, l.stream>>>(q, cur, keys, hist, sorted);
    sortCopyKernel<<<l.blocks, kBlock, 0, l.stream>>>(q, cur, sorted);
    return cudaGetLastError();
}
cudaError_t wfShadeDirect(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st,
                          Queues q, int cur, int strategy) {
    if (st.dirRec) shadeDirectKernel<true><<<l.blocks, kBlock, 0, l.stream>>>(sc, rc, st, q, cur, strategy);
    else shadeDirectKernel<false><<<l.blocks, kBlock, 0, l.stream>>>(sc, rc, st, q, cur, strategy);
    return cudaGetLastError();
}
cudaError_t wfTexturesDirect(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st, Queues q,
                             int cur) {
    textureDirectKernel<<<l.blocks, kBlock, 0, l.stream>>>(sc, rc, st, q, cur);
    return cudaGetLastError();
}
cudaError_t wfFilm(const WavefrontLaunch &l, const RenderConsts &rc, PathState st, Queues q, float4 *film,
                   float *hemiIntensity) {
    LAUNCH(filmKernel, l, rc, st, q, film, hemiIntensity)
}
cudaError_t wfResetCounters(const WavefrontLaunch &l, Queues q, int next) {
    resetCountersKernel<<<1, 1, 0, l.stream>>>(q, next);
    return cudaGetLastError();
}

cudaError_t buildSpatialLightGrid(const DeviceScene &sc, const float *halton, float *table, cudaStream_t stream) {
    const size_t nVox = (size_t)sc.voxels[0] * sc.voxels[1] * sc.voxels[2];
    float *contrib = nullptr;
    cudaError_t e = cudaMalloc((void **)&contrib, nVox * sc.nLights * sizeof(float));
    if (e != cudaSuccess) return e;
    lightGridAccumulate<<<148 * 8, 128, 0, stream>>>(sc, halton, contrib);
    lightGridFinalize<<<148 * 4, 128, 0, stream>>>(sc, contrib, table);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(contrib);
    return e;
}

}  // namespace iile
