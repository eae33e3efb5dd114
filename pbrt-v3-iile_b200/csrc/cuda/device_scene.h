// Device-resident scene: how the flattened scene is laid out in HBM.
//
//  nodes      (not uploaded)           the reference's binary LinearBVHNode[] (bvh.cpp:95-104) stays on the host;
//                                      the device form of the same tree is wideNodes
//  wideNodes  float4[8*nWide]          128 B each: two binary levels fused, child boxes SoA (traverse.cuh)
//  leafPrims  float4[3*nPrims]         BVH-leaf order.  Triangle slot: the three world-space
//                                      vertices pre-gathered; the spare .w lanes carry
//                                      {scene-order primitive index, kind/flags, sphere index}.
//                                      One 48 B contiguous read per primitive test.
//  prims      IilePrim[nPrims]         scene order (shading: vertex indices, material, light, flags)
//  P,N,UV     float[3n],[3n],[2n]      vertex pools (shading only)
//  spheres / materials / lights        small tables
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/iile_b200.h"

namespace iile {

// wide nodes the traversal kernels keep in shared memory: the top four levels, breadth first
constexpr int kTopStaged = 1 + 4 + 16 + 64;

struct DeviceScene {
    const float4 *nodes;      // always null: the binary LinearBVHNode array is not needed on the device
    const float4 *wideNodes;  // 8 float4 per 4-wide node (traverse.cuh), what the kernels walk
    const float4 *leafPrims;  // 3 float4 per primitive slot
    const IilePrim *prims;
    const float *P;
    const float *N;
    const float *UV;
    const IileSphere *spheres;
    const IileMaterial *materials;
    const IileLight *lights;
    // textures (texture.cuh): records, MIPMap pyramids in one texel pool, the per-material slots, and the EWA
    // filter's weight table (MIPMap::weightLut, mipmap.h:197-203)
    const IileTexture *textures;
    const IileImage *images;
    const float *texels;
    const IileMaterialTextures *materialTextures;
    const float *ewaWeights;
    // light sampling (lightdistrib.cpp): uniform / power CDF or spatial voxel grid
    const float *lightCdf;    // strategy-dependent table, see wavefront.cu
    // {next work item, finished blocks} of the persistent traversal kernels: zero between launches
    // (the last block of a launch resets them).  One pair per CUDA stream (scene_handle.h
    // deviceViewFor): launches on one stream are ordered, launches on different streams may overlap.
    uint32_t *castCounters;
    uint32_t nWide;
    uint32_t nNodes, nPrims, nSpheres, nMaterials, nLights;
    uint32_t nInfinite;  // lights of type IILE_LIGHT_INFINITE (escaped rays look at them)
    int lightStrategy;
    int voxels[3];
    float worldMin[3], worldMax[3];
};

// kind word stored in leafPrims[3*slot+1].w
constexpr uint32_t kLeafSphere = 0x80000000u;

}  // namespace iile
