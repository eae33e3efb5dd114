// Ray-cast kernels behind iile_intersect_closest / iile_intersect_any
// (BVHAccel::Intersect / IntersectP, src/accelerators/bvh.cpp:662-738) and the
// cfg-2 synthetic ray generator.
#include <algorithm>

#include "cast_kernels.h"
#include "libm_exact.cuh"
#include "pcg.cuh"
#include "traverse.cuh"

namespace iile {

namespace {

__device__ __forceinline__ void loadRay(const IileRay *rays, size_t i, RayCtx &r) {
    const float4 *p = reinterpret_cast<const float4 *>(rays + i);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    initRay(r, a.x, a.y, a.z, b.x, b.y, b.z, a.w);
}

// Ray batch in, hit records out: the Src of castPersistent for the C-ABI cast calls.
struct RayBatch {
    const IileRay *__restrict__ rays;
    IileHit *__restrict__ hits;
    uint8_t *__restrict__ occluded;
    const float4 *__restrict__ leafPrims;
    __device__ __forceinline__ bool load(uint32_t i, RayCtx &r) const {
        loadRay(rays, i, r);
        return true;
    }
    // 32-byte rays: a chunk of <= 256 rays is <= 64 lines of 128 bytes, two per lane
    __device__ __forceinline__ void prefetch(uint32_t first, uint32_t count, unsigned lane) const {
        const char *p = reinterpret_cast<const char *>(rays + first);
        const uint32_t bytes = count * (uint32_t)sizeof(IileRay);
        for (uint32_t off = lane * 128u; off < bytes; off += 32u * 128u) prefetchL2(p + off);
    }
    __device__ __forceinline__ void store(uint32_t i, const RayCtx &r, const HitRec &h, bool found) const {
        if (occluded) {
            occluded[i] = found ? 1 : 0;
            return;
        }
        int4 out;
        if (h.slot >= 0) {
            out.x = __float_as_int(__ldg(leafPrims + 3 * h.slot).w);  // scene-order primitive index
            out.y = __float_as_int(h.t);
            out.z = __float_as_int(h.b1);
            out.w = __float_as_int(h.b2);
        } else {
            out.x = -1;
            out.y = __float_as_int(r.tMax);
            out.z = 0;
            out.w = 0;
        }
        reinterpret_cast<int4 *>(hits)[i] = out;
    }
};

// The last block to finish puts the work counter back to zero for the next launch
// (atomicInc wraps the block counter itself).
__device__ __forceinline__ void releaseWorkCounter(uint32_t *counters) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicInc(&counters[1], gridDim.x - 1) == gridDim.x - 1) counters[0] = 0;
    }
}

__global__ void __launch_bounds__(kCastBlock, kCastMinBlocks) closestHitKernel(DeviceScene sc, const IileRay *__restrict__ rays,
                                                               uint32_t n, IileHit *__restrict__ hits) {
    RayBatch src{rays, hits, nullptr, sc.leafPrims};
    castPersistent<kCastClosest>(sc, src, n, sc.castCounters);
    releaseWorkCounter(sc.castCounters);
}

__global__ void __launch_bounds__(kCastBlock, kCastMinBlocks) anyHitKernel(DeviceScene sc, const IileRay *__restrict__ rays,
                                                           uint32_t n, uint8_t *__restrict__ occluded) {
    RayBatch src{rays, nullptr, occluded, sc.leafPrims};
    castPersistent<kCastAny>(sc, src, n, sc.castCounters);
    releaseWorkCounter(sc.castCounters);
}

// chunk c (2^14 rays) draws from RNG(c); ray j of the chunk uses
// draws 5j..5j+4 (ox oy oz u0 u1).  PCG32's O(log n) Advance (rng.h:116-140)
// lets every thread jump straight to its five draws.
__global__ void generateRaysKernel(float3 bmin, float3 bmax, uint64_t firstRay, size_t n,
                                   IileRay *__restrict__ rays) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t g = firstRay + i;
    Pcg32 rng;
    rng.setSequence(g >> 14);
    rng.advance((int64_t)(5 * (g & 16383)));
    float u[5];
    for (int k = 0; k < 5; ++k) u[k] = rng.uniformFloat();
    // Lerp(t, a, b) = (1-t)*a + t*b, pbrt.h:414
    float ox = (1 - u[0]) * bmin.x + u[0] * bmax.x;
    float oy = (1 - u[1]) * bmin.y + u[1] * bmax.y;
    float oz = (1 - u[2]) * bmin.z + u[2] * bmax.z;
    // UniformSampleSphere, sampling.cpp:98-103
    float z = 1 - 2 * u[3];
    float rr = sqrtf(fmaxf(0.f, 1 - z * z));
    float phi = 2 * 3.14159265358979323846f * u[4];
    float4 a = make_float4(ox, oy, oz, INFINITY);
    // std::cos / std::sin of a float on the oracle's side: the host's cosf / sinf (libm_exact.cuh), so that the rays
    // made here are the oracle's rays bit for bit
    float sp, cp;
    sincosGlibc(phi, &sp, &cp);
    float4 b = make_float4(rr * cp, rr * sp, z, 0.f);
    float4 *p = reinterpret_cast<float4 *>(rays + i);
    p[0] = a;
    p[1] = b;
}

}  // namespace

// Persistent grid: as many blocks as fit on the device at once (a multiple of the SM count).
template <typename K>
static int persistentBlocks(K kernel) {
    int dev = 0, sms = 148, perSm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kCastBlock, 0);
    return sms * (perSm > 0 ? perSm : 1);
}

cudaError_t launchClosestHit(const DeviceScene &sc, const IileRay *rays, size_t n, IileHit *hits,
                             cudaStream_t stream) {
    static const int blocks = persistentBlocks(closestHitKernel);
    for (size_t off = 0; off < n; off += (size_t)1 << 30) {
        const uint32_t m = (uint32_t)std::min<size_t>(n - off, (size_t)1 << 30);
        closestHitKernel<<<blocks, kCastBlock, 0, stream>>>(sc, rays + off, m, hits + off);
    }
    return cudaGetLastError();
}

cudaError_t launchAnyHit(const DeviceScene &sc, const IileRay *rays, size_t n, uint8_t *occluded,
                         cudaStream_t stream) {
    static const int blocks = persistentBlocks(anyHitKernel);
    for (size_t off = 0; off < n; off += (size_t)1 << 30) {
        const uint32_t m = (uint32_t)std::min<size_t>(n - off, (size_t)1 << 30);
        anyHitKernel<<<blocks, kCastBlock, 0, stream>>>(sc, rays + off, m, occluded + off);
    }
    return cudaGetLastError();
}

cudaError_t launchGenerateRays(const float bmin[3], const float bmax[3], uint64_t firstRay, size_t n,
                               IileRay *rays, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    unsigned blocks = (unsigned)((n + 255) / 256);
    generateRaysKernel<<<blocks, 256, 0, stream>>>(make_float3(bmin[0], bmin[1], bmin[2]),
                                                   make_float3(bmax[0], bmax[1], bmax[2]), firstRay,
                                                   n, rays);
    return cudaGetLastError();
}

}  // namespace iile
