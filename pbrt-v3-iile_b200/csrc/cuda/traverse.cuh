// Stack-based near-first BVH traversal, one ray per thread.
//
// Visits nodes in exactly the order of BVHAccel::Intersect / IntersectP
// (src/accelerators/bvh.cpp:662-738): test the node's box with the CURRENT
// ray.tMax; leaf -> test its primitives in array order, each accepted hit
// shrinking tMax (GeometricPrimitive::Intersect, primitive.cpp:116-130);
// interior -> descend into the child on the near side of the split axis and
// push the other.  A later primitive with t == tMax replaces the earlier one
// (triangle.cpp:258-261 rejects only strictly greater), so the visiting order
// is part of the result and is kept.
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

template <bool ANY_HIT>
__device__ __forceinline__ bool traverse(const DeviceScene &sc, RayCtx &r, HitRec *hit) {
    int stack[64];
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
                    const int slot = offset + i;
                    const float4 v0 = __ldg(leaf + 3 * slot);
                    const float4 v1 = __ldg(leaf + 3 * slot + 1);
                    const float4 v2 = __ldg(leaf + 3 * slot + 2);
                    float t, b0, b1, b2;
                    bool h;
                    if (__float_as_uint(v1.w) & kLeafSphere)
                        h = sphereTest(sc.spheres[__float_as_int(v2.w)], r, r.tMax, &t, &b1, &b2),
                        b0 = 0.f;
                    else
                        h = triangleTest(r, r.tMax, make_float3(v0.x, v0.y, v0.z),
                                         make_float3(v1.x, v1.y, v1.z),
                                         make_float3(v2.x, v2.y, v2.z), &t, &b0, &b1, &b2);
                    if (h) {
                        if (ANY_HIT) return true;
                        found = true;
                        r.tMax = t;
                        hit->slot = slot;
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
