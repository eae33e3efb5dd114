// BVH traversal: persistent warps, "while-while" loop structure, dynamic ray fetch.
//
// Per ray the nodes and primitives are visited in exactly the order of
// BVHAccel::Intersect / IntersectP (src/accelerators/bvh.cpp:662-738): test the
// node's box with the CURRENT ray.tMax; leaf -> test its primitives in array
// order, each accepted hit shrinking tMax (GeometricPrimitive::Intersect,
// primitive.cpp:116-130); interior -> descend into the child on the near side
// of the split axis and push the other.  A later primitive with t == tMax
// replaces the earlier one (triangle.cpp:258-261 rejects only strictly
// greater), so the visiting order is part of the result and is kept: the
// restructuring below changes how the 32 lanes of a warp are scheduled, never
// the sequence of tests one ray sees.
//
// Scheduling (incoherent rays leave a one-ray-per-thread loop at ~4 of 32
// lanes busy, see profiles/r1a_*):
//   * node expansion and primitive tests are separate phases of one loop; each iteration
//     runs the phase more lanes are waiting for (see castPersistent).
//   * persistent warps + dynamic fetch: a warp owns a chunk of 32..256 rays
//     taken from a global counter; a lane whose ray has terminated is refilled
//     from the chunk as soon as kRefill lanes are idle, so long rays do not keep
//     a warp at low occupancy.
//   * two levels of the binary tree are fused into one 128-byte 4-wide node (one cache
//     line, eight 16-byte loads, four independent slab tests): a third of the dependent
//     fetches, and culled stack entries cost a compare instead of a node fetch.
//
// Memory: a node is two 16-byte loads (read-only path), a primitive slot three.
#pragma once

#include "device_scene.h"
#include "intersect.cuh"
#include "sphere.cuh"

namespace iile {

struct HitRec {
    int slot;  // BVH-order primitive slot, -1 = miss
    float t, b0, b1, b2;
};

__device__ __forceinline__ void prefetchL2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ void prefetchL1(const void *p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

constexpr int kCastBlock = 128;   // threads per block of every traversal kernel
#ifndef IILE_CAST_MINBLOCKS
#define IILE_CAST_MINBLOCKS 7
#endif
constexpr int kCastMinBlocks = IILE_CAST_MINBLOCKS;  // resident blocks per SM asked of ptxas (7: 72 registers)
constexpr uint32_t kChunkMax = 256;  // most rays a warp takes from the global counter at a time
constexpr uint32_t kChunkMin = 32;
#ifndef IILE_GUIDED_CHUNKS
#define IILE_GUIDED_CHUNKS 1
#endif
#ifndef IILE_GUIDED_MIN
#define IILE_GUIDED_MIN 32   // smallest guided chunk
#endif
#ifndef IILE_GUIDED_DIV
#define IILE_GUIDED_DIV 2    // a chunk is the remaining rays / (warps of the grid * this)
#endif
#ifndef IILE_EXPAND_CULL
#define IILE_EXPAND_CULL 0   // exact; costs <= 1 % against the pre-cull (profiles/r3e_ab_expand_cull_*.txt)
#endif
#ifndef IILE_KREFILL
#define IILE_KREFILL 16
#endif
constexpr int kRefill = IILE_KREFILL;  // refill idle lanes once this many are idle

// One primitive test of a leaf (triangle or sphere slot).
__device__ __forceinline__ bool primitiveTest(const DeviceScene &sc, const float4 *__restrict__ leaf, int slot,
                                              const RayCtx &r, float *t, float *b0, float *b1, float *b2) {
    const float4 v0 = __ldg(leaf + 3 * slot);
    const float4 v1 = __ldg(leaf + 3 * slot + 1);
    const float4 v2 = __ldg(leaf + 3 * slot + 2);
    if (__float_as_uint(v1.w) & kLeafSphere) {
        *b0 = 0.f;
        return sphereTest(sc.spheres[__float_as_int(v2.w)], r, r.tMax, t, b1, b2);
    }
    return triangleTest(r, r.tMax, make_float3(v0.x, v0.y, v0.z), make_float3(v1.x, v1.y, v1.z),
                        make_float3(v2.x, v2.y, v2.z), t, b0, b1, b2);
}

// ---- 4-wide node: two levels of the reference's binary tree in one 128-byte record ----------
//
// Wide node of binary interior node N (children A = N+1, B = N.secondChildOffset):
//   slots 0,1 = children of A (or A itself in slot 0 when A is a leaf), slots 2,3 = those of B.
//   float4 [0] minx[4] [1] maxx[4] [2] miny[4] [3] maxy[4] [4] minz[4] [5] maxz[4]
//   uint4  [6] ref[4]   leaf: 1<<31 | nPrims<<23 | first primitive slot;  interior: wide index
//   uint   [7].x meta   axis(N) | axis(A)<<2 | axis(B)<<4 | validSlots<<8
// An unused slot carries the box [+FLT_MAX, -FLT_MAX]: its far plane is behind every ray
// (tMax = -big or -inf, never NaN), so the slab test rejects it without a validity check.
// Visiting order of the slots = the reference's near-first order: group A/B by the sign of
// d[axis(N)], inside a group by the sign of d[axis(A|B)].  The boxes of A and B themselves
// are not tested: a child box that passes lies inside its parent's, whose test (same
// monotone float arithmetic, same or larger tMax) would have passed too, so the set and
// order of leaves visited is unchanged.  This holds for NaN slabs as well (ADVICE r1 asked): a ray with
// an infinite 1/d component whose origin lies exactly on a plane of A's box gets 0 * inf = NaN there and
// the reference's comparisons cull A (geometry.h:1420-1437) — but every child of A then fails too: on
// that axis the ray never leaves the plane, a child box touching the plane meets the same NaN in the same
// place, and a child strictly inside gets +inf as its entry distance (or -inf as its exit distance) and is
// rejected.  tests/test_cast_gpu.py::test_axis_aligned_rays_on_box_planes casts 24,576 axis-aligned rays
// from integer-aligned origins through a box-aligned scene: 0 differ from the reference.  (Three ways of
// repeating the parents' tests anyway were built and measured before this was settled; they cost the
// loop 6-20 % and are gone: profiles/r7_ab_nan_variants.txt.)  Each pushed entry keeps its entry distance and is
// re-checked against the CURRENT tMax when popped — exactly the `tMin < ray.tMax` the
// reference evaluates when it reaches that node (geometry.h:1437).
constexpr uint32_t kWideLeaf = 0x80000000u;
constexpr uint32_t kWideSentinel = 0xffffffffu;
// IILE_TOP_SMEM=1 copies the first kTopStaged wide nodes (the tree's top four levels, breadth first:
// c_api.cu) into shared memory in every block of a traversal kernel, so that the nodes every ray
// walks are read from shared memory instead of L1 (85 nodes = 10.6 KB per block, 7 blocks per SM).
// Measured on B200 it LOSES 2-5 % on every workload (profiles/r3c_ab_top_smem_*.txt: cfg2 11.29 ->
// 10.79 Grays/s, cfg1 431 -> 418 Msamples/s): those lines already sit in L1 (87 % L1 hit rate in
// the render's cast kernel, profiles/r2c), the generic loads give up the read-only path, and
// 76 KB of shared memory per SM comes out of the same SRAM the L1 uses.  Off by default; the
// breadth-first numbering of the top is kept (it costs nothing and keeps those lines together).
#ifndef IILE_TOP_SMEM
#define IILE_TOP_SMEM 0
#endif
constexpr int kWideStack = 96;  // entries; uploadScene checks the tree against this bound
// IILE_TOS_CACHE=1 keeps a copy of the stack's top entry in two registers (the stack in local memory
// stays complete: write-through).  A pop takes the register copy and issues the load of the entry
// below it at once, so that load's latency runs behind the expansion / primitive test that follows
// instead of in front of the next pop (ncu: 13.5 % of the closest-hit kernel's stall samples sat on
// the pop's local load, profiles/r5y_closestHit_source_hot.txt).  Control flow and results unchanged
// (bit-exact on the cast and watertight parity tests) — but measured on B200 it LOSES 1-3 %
// (profiles/r6b_ab_tos_cache.txt: cfg2 closest-hit 1.575 -> 1.595 ms, terrain 20.8 -> 20.4 Grays/s,
// cfg1 451 -> 441, cfg3 206 -> 199 Msamples/s): the four predicated register moves per expansion
// cost more issue slots than the hidden latency returns at 28 warps per SM.  Off by default.
#ifndef IILE_TOS_CACHE
#define IILE_TOS_CACHE 0
#endif
// IILE_SMEM_STACK=N keeps the bottom N entries of every lane's traversal stack in shared memory
// (sStack[entry][thread]: conflict-free 8-byte accesses) and only the entries above them in local memory —
// with N = 16 or 24 nearly every push and pop of cfg1 / cfg2 / cfg3 / cfg5 stays out of the L1's
// local-memory traffic (VERDICT r1 next #5 asked for the experiment).  0 = the stack lives in local memory.
#ifndef IILE_SMEM_STACK
#define IILE_SMEM_STACK 0
#endif
// Which phase an iteration runs: the test phase when nHold * HOLD_WEIGHT > nWalk * WALK_WEIGHT.
// The plain majority (1 : 1) is the measured optimum on B200 (profiles/r6d_ab_phase_weights.txt:
// cfg2 11.73 Grays/s; 2:1 11.52, 1:2 11.46, 3:2 11.67; cfg1 450 / 449 / 438 / 453 Msamples/s).
// Prefetching the children an expansion pushes for later (prefetch.global.L2 / .L1 of their node or
// first primitive line) was measured on B200 and dropped: it loses 4-5 % on the scene that does not fit
// the L2 (terrain casts 19.7 -> 18.8 Grays/s, cfg5 485 -> 462 Msamples/s) and the dormant branch alone
// costs 1 % on the L2-resident ones (profiles/r7_ab_nan_variants.txt).
#ifndef IILE_PHASE_HOLD_WEIGHT
#define IILE_PHASE_HOLD_WEIGHT 1
#endif
#ifndef IILE_PHASE_WALK_WEIGHT
#define IILE_PHASE_WALK_WEIGHT 1
#endif

// geometry.h:1412-1438 on already sign-selected planes (near / far per axis).  Returns the
// slab result WITHOUT the `tMin < ray.tMax` term; *tMinOut is the entry distance.
__device__ __forceinline__ bool slabPlanes(float nearX, float farX, float nearY, float farY, float nearZ,
                                           float farZ, const RayCtx &r, float *tMinOut) {
    const float k = 1 + 2 * gammaF(3);
    float tMin = (nearX - r.ox) * r.ix;
    float tMax = (farX - r.ox) * r.ix;
    float tyMin = (nearY - r.oy) * r.iy;
    float tyMax = (farY - r.oy) * r.iy;
    tMax *= k;
    tyMax *= k;
    bool ok = !(tMin > tyMax || tyMin > tMax);
    if (tyMin > tMin) tMin = tyMin;
    if (tyMax < tMax) tMax = tyMax;
    float tzMin = (nearZ - r.oz) * r.iz;
    float tzMax = (farZ - r.oz) * r.iz;
    tzMax *= k;
    ok = ok && !(tMin > tzMax || tzMin > tMax);
    if (tzMin > tMin) tMin = tzMin;
    if (tzMax < tMax) tMax = tzMax;
    *tMinOut = tMin;
    return ok && tMax > 0;
}

// MODE of a traversal: kCastClosest / kCastAny fix the query for the whole launch; kCastMixed lets every
// ray carry its own (Src::load(i, r, any) sets it) so that one launch serves several ray queues.
enum : int { kCastClosest = 0, kCastAny = 1, kCastMixed = 2 };

// Src: `bool load(uint32_t i, RayCtx &r)` (false = nothing to trace for item i),
//      `void store(uint32_t i, const RayCtx &r, const HitRec &h, bool found)` and
//      `void prefetch(uint32_t first, uint32_t count, unsigned lane)` (a hint: the warp just took
//      items [first, first+count) and will load them over the next iterations).
// All 32 lanes of every warp of the grid must call this together.  `workCounter`
// must be zero at kernel start.
//
// One loop, three phases per iteration:
//   pop     a lane that needs its next node takes entries off its stack until one is still
//           in front of tMax: a wide node (expanded below) or a leaf (held).
//   expand  every lane with a wide node fetches its 128 bytes, runs the four slab tests and
//           pushes the passing children in reverse visiting order.
//   test    a lane holding a leaf must stop there — the leaf's hits shrink tMax, which later
//           box checks read — and runs one primitive test per test phase.
// An iteration runs either the expand or the test phase, whichever more lanes wait for.
// Finished lanes keep their result until kRefill lanes are idle; then results are stored and
// new rays loaded in one go.
//
// MODE: kCastClosest / kCastAny fix the query for the whole launch; kCastMixed lets every ray
// carry its own (Src::load(i, r, any) sets it) so that one launch serves several ray queues.

template <int MODE, class Src>
__device__ __forceinline__ void castPersistent(const DeviceScene &sc, Src &src, uint32_t n,
                                               uint32_t *workCounter) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const float4 *__restrict__ wide = sc.wideNodes;
    const float4 *__restrict__ leaf = sc.leafPrims;
#if IILE_SMEM_STACK
    __shared__ uint2 sStack[IILE_SMEM_STACK][kCastBlock];
    uint2 stackHi[kWideStack - IILE_SMEM_STACK];
#define STACK_PUT(i, v)                                              \
    do {                                                             \
        const int i__ = (i);                                         \
        if (i__ < IILE_SMEM_STACK) sStack[i__][threadIdx.x] = (v);   \
        else stackHi[i__ - IILE_SMEM_STACK] = (v);                   \
    } while (0)
#define STACK_GET(i) ((i) < IILE_SMEM_STACK ? sStack[(i)][threadIdx.x] : stackHi[(i) - IILE_SMEM_STACK])
#else
    uint2 stack[kWideStack];
#define STACK_PUT(i, v) stack[(i)] = (v)
#define STACK_GET(i) stack[(i)]
#endif
#if IILE_TOP_SMEM
    __shared__ float4 sTop[kTopStaged * 8];
    const int topCount = min((int)sc.nWide, kTopStaged);
    for (int k = threadIdx.x; k < topCount * 8; k += blockDim.x) sTop[k] = __ldg(wide + k);
    __syncthreads();
#endif
    uint32_t chunkNext = 0, chunkEnd = 0;  // warp-uniform
    // chunk size: large batches amortise the global atomic over 256 rays; small ones (late bounces)
    // are cut finer so that every warp of the grid gets work instead of a few warps grinding
    // through 256 rays each (a lone warp needs ~100 us for that)
    const uint32_t kChunk = min(kChunkMax, max(kChunkMin, (n / (gridDim.x * (blockDim.x >> 5) * 4u)) & ~31u));
    bool exhausted = false;                // warp-uniform: the global counter ran past n
    RayCtx r;
    HitRec hit;
    uint32_t item = 0;
    int top = 0;          // next free stack entry
#if IILE_TOS_CACHE
    uint2 tos = make_uint2(kWideSentinel, 0u);  // == stack[top - 1] whenever top > 0
#endif
    int cur = -1;         // wide node to expand, -1 none
    int leafOff = 0, leafCnt = 0;
    bool needPop = false, found = false, pending = false;
    bool anyHit = MODE == kCastAny;  // per lane; only kCastMixed changes it
    hit.slot = -1;
    hit.t = hit.b0 = hit.b1 = hit.b2 = 0.f;
    for (;;) {
        // ---- retire finished rays and refill idle lanes from the warp's chunk -----------------
        const bool idle = cur < 0 && leafCnt == 0 && !needPop;
        const unsigned idleMask = __ballot_sync(FULL, idle);
        const int nIdle = __popc(idleMask);
        if (nIdle >= kRefill) {
            if (idle && pending) {
                src.store(item, r, hit, found);
                pending = false;
            }
            if (chunkNext == chunkEnd && !exhausted) {
                uint32_t base = 0, take = kChunk;
                if (lane == 0) {
#if IILE_GUIDED_CHUNKS
                    // guided self-scheduling: chunks shrink as the counter nears n, so that the warps of
                    // the grid finish together instead of idling behind the last 256-ray chunks
                    // (profiles/r2c: 9 % of castKernel's stall samples sat on the closing barrier)
                    const uint32_t seen = *(volatile uint32_t *)workCounter;
                    const uint32_t rem = seen < n ? n - seen : 0u;
                    // the wavefront's mixed launches (long, uneven rays) taper to 32-ray chunks over the last
                    // 1/8 of the work; the uniform casts of the ray-cast API to 64 over the last 1/16 — their
                    // rays are short enough for the counter's atomics to show (A/B: profiles/r2j_ab_*.txt)
                    constexpr uint32_t gMin = MODE == kCastMixed ? (uint32_t)IILE_GUIDED_MIN : 2u * IILE_GUIDED_MIN;
                    constexpr uint32_t gDiv = MODE == kCastMixed ? (uint32_t)IILE_GUIDED_DIV : (uint32_t)IILE_GUIDED_DIV / 2u;
                    take = min(kChunk, max(gMin, (rem / (gridDim.x * (blockDim.x >> 5) * (gDiv ? gDiv : 1u))) & ~31u));
#endif
                    base = atomicAdd(workCounter, take);
                }
                base = __shfl_sync(FULL, base, 0);
                take = __shfl_sync(FULL, take, 0);
                if (base >= n) exhausted = true;
                else {
                    chunkNext = base;
                    chunkEnd = min(base + take, n);
                    src.prefetch(chunkNext, chunkEnd - chunkNext, lane);  // the chunk's rays on their way to L2
                }
            }
            const uint32_t avail = chunkEnd - chunkNext;
            if (avail) {
                const uint32_t rank = __popc(idleMask & ((1u << lane) - 1));
                if (idle && rank < avail) {
                    item = chunkNext + rank;
                    bool ok;
                    if constexpr (MODE == kCastMixed) ok = src.load(item, r, anyHit);
                    else ok = src.load(item, r);
                    if (ok) {
                        hit.slot = -1;
                        found = false;
                        pending = true;
                        STACK_PUT(0, make_uint2(kWideSentinel, 0u));
                        top = 1;
                        cur = 0;

#if IILE_TOS_CACHE
                        tos = make_uint2(kWideSentinel, 0u);
#endif
                    }
                }
                chunkNext += min(avail, (uint32_t)nIdle);
            }
            if (nIdle == 32 && !__any_sync(FULL, cur >= 0)) {
                if (exhausted) break;
                continue;  // chunk ran dry (or every item was empty): take the next one
            }
        }
        // ---- pop ---------------------------------------------------------------------------------
        if (needPop) {
            needPop = false;
            for (;;) {
#if IILE_TOS_CACHE
                const uint2 e = tos;
                --top;
                if (e.x == kWideSentinel) break;  // cur stays -1: the ray is finished
                tos = STACK_GET(top - 1);         // needed at the next pop only
#else
                --top;
                const uint2 e = STACK_GET(top);
                if (e.x == kWideSentinel) break;  // cur stays -1: the ray is finished
#endif
                if (__uint_as_float(e.y) < r.tMax) {
                    if (e.x & kWideLeaf) {
                        leafOff = (int)(e.x & 0x7fffffu);
                        leafCnt = (int)((e.x >> 23) & 0xffu);
                    } else
                        cur = (int)e.x;
                    break;
                }
            }
        }
        // ---- which phase runs this iteration: the one more lanes are waiting for -----------------
        const bool holding = leafCnt > 0;
        const int nHold = __popc(__ballot_sync(FULL, holding));
        const int nWalk = __popc(__ballot_sync(FULL, cur >= 0));
        const bool testPhase = nHold * IILE_PHASE_HOLD_WEIGHT > nWalk * IILE_PHASE_WALK_WEIGHT;
        // ---- expand ------------------------------------------------------------------------------
        if (!testPhase && cur >= 0) {
            const int w = cur * 8;
#if IILE_TOP_SMEM
            // one generic load per field: the address decides between shared memory and L1 / L2
            const float4 *nb = cur < topCount ? sTop + w : wide + w;
            const float4 nX = nb[r.nx], fX = nb[r.nx ^ 1];
            const float4 nY = nb[2 + r.ny], fY = nb[2 + (r.ny ^ 1)];
            const float4 nZ = nb[4 + r.nz], fZ = nb[4 + (r.nz ^ 1)];
            const uint4 ref = *reinterpret_cast<const uint4 *>(nb + 6);
            const uint32_t meta = *reinterpret_cast<const uint32_t *>(nb + 7);
#else
            const float4 nX = __ldg(wide + w + r.nx), fX = __ldg(wide + w + (r.nx ^ 1));
            const float4 nY = __ldg(wide + w + 2 + r.ny), fY = __ldg(wide + w + 2 + (r.ny ^ 1));
            const float4 nZ = __ldg(wide + w + 4 + r.nz), fZ = __ldg(wide + w + 4 + (r.nz ^ 1));
            const uint4 ref = __ldg(reinterpret_cast<const uint4 *>(wide + w + 6));
            const uint32_t meta = __ldg(reinterpret_cast<const uint32_t *>(wide + w + 7));
#endif
            float t0, t1, t2, t3;
            // The binding `tMin < ray.tMax` of geometry.h:1437 is the one made when an entry is popped
            // (the moment the reference tests that box).  Here, two levels earlier, children are only
            // pre-culled — and for closest-hit rays not against tMax itself: Triangle::Intersect accepts
            // tScaled <= tMax * det and then stores t = tScaled / det, which can round to one ulp ABOVE the
            // old tMax (triangle.cpp:247-270), so ray.tMax is not monotone and a box that fails against
            // today's tMax may pass when the reference reaches it (found by the Triangle.Watertight rays
            // aimed at shared vertices).  IILE_EXPAND_CULL: 0 = no pre-cull for closest-hit rays (exact),
            // 1 = pre-cull against tMax (the pre-fix behaviour), 2 = against tMax * (1 + 1e-4).
            float cullT = r.tMax;
            if (!(MODE == kCastAny || (MODE == kCastMixed && anyHit))) {
#if IILE_EXPAND_CULL == 0
                cullT = INFINITY;
#elif IILE_EXPAND_CULL == 2
                cullT = r.tMax * 1.0001f;
#endif
            }
            const bool p0 = slabPlanes(nX.x, fX.x, nY.x, fY.x, nZ.x, fZ.x, r, &t0) && t0 < cullT;
            const bool p1 = slabPlanes(nX.y, fX.y, nY.y, fY.y, nZ.y, fZ.y, r, &t1) && t1 < cullT;
            const bool p2 = slabPlanes(nX.z, fX.z, nY.z, fY.z, nZ.z, fZ.z, r, &t2) && t2 < cullT;
            const bool p3 = slabPlanes(nX.w, fX.w, nY.w, fY.w, nZ.w, fZ.w, r, &t3) && t3 < cullT;
            // visiting position of each slot (0 = first)
            const uint32_t sN = (r.negMask >> (meta & 3u)) & 1u;
            const uint32_t sA = (r.negMask >> ((meta >> 2) & 3u)) & 1u;
            const uint32_t sB = (r.negMask >> ((meta >> 4) & 3u)) & 1u;
            const uint32_t k0 = (sN << 1) | sA, k1 = k0 ^ 1u;
            const uint32_t k2 = ((sN ^ 1u) << 1) | sB, k3 = k2 ^ 1u;
            const uint32_t ordered = ((uint32_t)p0 << k0) | ((uint32_t)p1 << k1) | ((uint32_t)p2 << k2) |
                                     ((uint32_t)p3 << k3);
            // first visited ends on top: position = top + #passing slots visited later
            if (p0) STACK_PUT(top + __popc(ordered >> (k0 + 1)), make_uint2(ref.x, __float_as_uint(t0)));
            if (p1) STACK_PUT(top + __popc(ordered >> (k1 + 1)), make_uint2(ref.y, __float_as_uint(t1)));
            if (p2) STACK_PUT(top + __popc(ordered >> (k2 + 1)), make_uint2(ref.z, __float_as_uint(t2)));
            if (p3) STACK_PUT(top + __popc(ordered >> (k3 + 1)), make_uint2(ref.w, __float_as_uint(t3)));
#if IILE_TOS_CACHE
            {
                // the first visited passing child (the lowest set bit of `ordered`) is the new top
                const uint32_t first = (uint32_t)__ffs((int)ordered) - 1u;  // no child passes: 0xffffffff
                if (p0 && k0 == first) tos = make_uint2(ref.x, __float_as_uint(t0));
                if (p1 && k1 == first) tos = make_uint2(ref.y, __float_as_uint(t1));
                if (p2 && k2 == first) tos = make_uint2(ref.z, __float_as_uint(t2));
                if (p3 && k3 == first) tos = make_uint2(ref.w, __float_as_uint(t3));
            }
#endif
            top += __popc(ordered);
            cur = -1;
            needPop = true;
        }
        // ---- primitive test ------------------------------------------------------------------------
        {
            if (testPhase && holding) {
                float t, b0, b1, b2;
                if (primitiveTest(sc, leaf, leafOff, r, &t, &b0, &b1, &b2)) {
                    found = true;
                    if (MODE == kCastAny || (MODE == kCastMixed && anyHit)) {
                        leafCnt = 1;  // drop the rest of the leaf and the stack: the ray is done
                        top = 1;
#if IILE_TOS_CACHE
                        tos = make_uint2(kWideSentinel, 0u);
#endif
                    } else {
                        r.tMax = t;
                        hit.slot = leafOff;
                        hit.t = t;
                        hit.b0 = b0;
                        hit.b1 = b1;
                        hit.b2 = b2;
                    }
                }
                ++leafOff;
                if (--leafCnt == 0) needPop = true;
            }
        }
    }
#undef STACK_PUT
#undef STACK_GET
}

}  // namespace iile
