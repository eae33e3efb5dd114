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
//   * node steps and primitive tests are separate phases of one loop; primitive tests are
//     batched until kLeafBatch lanes wait at a leaf (see castPersistent).
//   * persistent warps + dynamic fetch: a warp owns a chunk of kChunk rays
//     taken from a global counter; a lane whose ray has terminated is refilled
//     from the chunk as soon as kRefill lanes are idle, so long rays do not keep
//     a warp at low occupancy.
//   * the traversal stack lives in shared memory, [level][thread], so pushes
//     and pops are conflict-free LDS/STS instead of local-memory traffic.
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

constexpr int kStackDepth = 64;   // nodesToVisit[64], bvh.cpp:670
constexpr int kCastBlock = 128;   // threads per block of every traversal kernel
constexpr uint32_t kChunk = 256;  // rays a warp takes from the global counter at a time
constexpr int kRefill = 8;        // refill idle lanes once this many are idle
constexpr int kLeafBatch = 16;    // run a primitive test once this many lanes hold a leaf
constexpr size_t kCastSmem = (size_t)(kStackDepth + 1) * kCastBlock * sizeof(int);  // + sentinel level

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

// 32-bit shared-memory accesses by address (keeps the stack off the generic-address path)
__device__ __forceinline__ void stsInt(uint32_t addr, int v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ldsInt(uint32_t addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// Src: `bool load(uint32_t i, RayCtx &r)` (false = nothing to trace for item i) and
//      `void store(uint32_t i, const RayCtx &r, const HitRec &h, bool found)`.
// All 32 lanes of every warp of the grid must call this together.  `workCounter`
// must be zero at kernel start; `stackMem` is the block's shared stack,
// kStackDepth x kCastBlock ints, [level][thread].
//
// One loop, two phases per iteration.  Node phase: every lane that is walking interior
// nodes takes ONE step (fetch, slab test, push / pop).  A lane that reaches a leaf must
// stop there — the leaf's hits shrink tMax, which the following box tests read — and
// waits.  Primitive phase: once kLeafBatch lanes hold a leaf (or nobody can step), the
// holders run one primitive test together.  This keeps most lanes busy in both phases
// where a per-lane "walk to the next leaf" loop idles on the slowest lane.
template <bool ANY_HIT, class Src>
__device__ __forceinline__ void castPersistent(const DeviceScene &sc, Src &src, uint32_t n,
                                               uint32_t *workCounter, int *stackMem) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const float4 *__restrict__ nodes = sc.nodes;
    const float4 *__restrict__ leaf = sc.leafPrims;
    const uint32_t stackBase = (uint32_t)__cvta_generic_to_shared(stackMem) + threadIdx.x * 4u;
    constexpr uint32_t kLevel = kCastBlock * 4u;  // bytes between stack levels
    uint32_t chunkNext = 0, chunkEnd = 0;  // warp-uniform
    bool exhausted = false;                // warp-uniform: the global counter ran past n
    bool active = false;
    RayCtx r;
    HitRec hit;
    uint32_t item = 0, top = stackBase;
    int cur = -1, leafOff = 0, leafCnt = 0;
    bool found = false;
    hit.slot = -1;
    hit.t = hit.b0 = hit.b1 = hit.b2 = 0.f;
    for (;;) {
        // ---- refill idle lanes from the warp's chunk --------------------------------------
        const unsigned idleMask = __ballot_sync(FULL, !active);
        const int nIdle = __popc(idleMask);
        if (nIdle >= kRefill) {
            if (chunkNext == chunkEnd && !exhausted) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(workCounter, kChunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) exhausted = true;
                else {
                    chunkNext = base;
                    chunkEnd = min(base + kChunk, n);
                }
            }
            const uint32_t avail = chunkEnd - chunkNext;
            if (avail) {
                const uint32_t rank = __popc(idleMask & ((1u << lane) - 1));
                if (!active && rank < avail) {
                    item = chunkNext + rank;
                    if (src.load(item, r)) {
                        hit.slot = -1;
                        found = false;
                        cur = 0;
                        leafCnt = 0;
                        stsInt(stackBase, -1);  // sentinel: popping it ends the ray
                        top = stackBase + kLevel;
                        active = true;
                    }
                }
                chunkNext += min(avail, (uint32_t)nIdle);
            }
            if (nIdle == 32 && !__any_sync(FULL, active)) {
                if (exhausted) break;
                continue;  // chunk ran dry (or every item was empty): take the next one
            }
        }
        // ---- node phase: one step for every lane that is between leaves ---------------------
        const bool walking = active && leafCnt == 0;
        if (walking) {
            const float4 n0 = __ldg(nodes + 2 * cur);
            const float4 n1 = __ldg(nodes + 2 * cur + 1);
            float tEntry;
            const bool hitBox = slabTest(n0, n1, r, &tEntry) && tEntry < r.tMax;
            const uint32_t meta = __float_as_uint(n1.w);
            const int nPrims = meta & 0xffff;
            const int offset = __float_as_int(n1.z);
            if (hitBox && nPrims == 0) {
                const int neg = (r.negMask >> ((meta >> 16) & 3)) & 1;
                stsInt(top, neg ? cur + 1 : offset);  // far child
                top += kLevel;
                cur = neg ? offset : cur + 1;         // near child
            } else {
                if (hitBox) {
                    leafOff = offset;
                    leafCnt = nPrims;
                }
                top -= kLevel;
                cur = ldsInt(top);
            }
        }
        // ---- primitive phase ----------------------------------------------------------------
        const bool holding = active && leafCnt > 0;
        const unsigned holdMask = __ballot_sync(FULL, holding);
        if (holdMask) {
            const unsigned walkMask = __ballot_sync(FULL, active && leafCnt == 0 && cur >= 0);
            if (__popc(holdMask) >= kLeafBatch || walkMask == 0) {
                if (holding) {
                    float t, b0, b1, b2;
                    if (primitiveTest(sc, leaf, leafOff, r, &t, &b0, &b1, &b2)) {
                        found = true;
                        if (ANY_HIT) {
                            cur = -1;
                            leafCnt = 1;
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
                    --leafCnt;
                }
            }
        }
        // ---- retire finished rays -----------------------------------------------------------------
        if (active && cur < 0 && leafCnt == 0) {
            src.store(item, r, hit, found);
            active = false;
        }
    }
}

// One ray, one thread, private stack: used where a single extra ray is traced from inside
// another kernel.  Same visiting order as above.
template <bool ANY_HIT>
__device__ __forceinline__ bool traverse(const DeviceScene &sc, RayCtx &r, HitRec *hit) {
    int stack[kStackDepth];
    int sp = 0;
    int cur = 0;
    bool found = false;
    if (!ANY_HIT) hit->slot = -1;
    const float4 *__restrict__ nodes = sc.nodes;
    const float4 *__restrict__ leaf = sc.leafPrims;
    while (true) {
        const float4 n0 = __ldg(nodes + 2 * cur);
        const float4 n1 = __ldg(nodes + 2 * cur + 1);
        float tEntry;
        if (slabTest(n0, n1, r, &tEntry) && tEntry < r.tMax) {
            const uint32_t meta = __float_as_uint(n1.w);
            const int nPrims = meta & 0xffff;
            const int offset = __float_as_int(n1.z);
            if (nPrims > 0) {
                for (int i = 0; i < nPrims; ++i) {
                    float t, b0, b1, b2;
                    if (primitiveTest(sc, leaf, offset + i, r, &t, &b0, &b1, &b2)) {
                        if (ANY_HIT) return true;
                        found = true;
                        r.tMax = t;
                        hit->slot = offset + i;
                        hit->t = t;
                        hit->b0 = b0;
                        hit->b1 = b1;
                        hit->b2 = b2;
                    }
                }
                if (sp == 0) break;
                cur = stack[--sp];
            } else {
                const int axis = (meta >> 16) & 0xff;
                const int neg = axis == 0 ? r.nx : (axis == 1 ? r.ny : r.nz);
                if (neg) {
                    stack[sp++] = cur + 1;
                    cur = offset;
                } else {
                    stack[sp++] = offset;
                    cur = cur + 1;
                }
            }
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    return found;
}

}  // namespace iile
