// Wavefront path integrator: host-visible types and launch wrappers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_scene.h"
#include "halton.cuh"

namespace iile {

// Everything the camera / film / integrator kernels need, passed by value.
struct RenderConsts {
    float rasterToCamera[16];
    float cameraToWorld[16];
    float worldToCamera[16];  // texture.cuh: the lens point of a depth-of-field ray is recovered through it
    float lensRadius, focalDistance;
    int cameraType;          // IILE_CAMERA_*
    float fullRes[2];        // film->fullResolution (environment camera)
    float filterTable[256];
    float filterRadius[2], invFilterRadius[2];
    float maxSampleLuminance;
    int filmMonitor;  // IILE_FILTER_MONITOR: a sample lands on the pixel it was drawn for
    int sampleBounds[4];   // x0 y0 x1 y1
    int pixelBounds[4];    // x0 y0 x1 y1 (integrator pixelbounds ∩ sample bounds)
    int cropped[4];        // x0 y0 x1 y1 film pixels
    int spp;               // samples per pixel of the whole render (sequence index space)
    int sampleBegin, sampleEnd;
    int tilesX, tilesY, tileBegin, tileEnd;
    int maxDepth;
    float rrThreshold;
    // hemispheric mode (IISPTdIntegrator::RenderView): camera = per-item position/axis
    int hemiMode, hemiSize;
    // Sampler "halton": tables == nullptr selects the random sampler (one PCG32 stream per path)
    HaltonConsts halton;
    // GlobalSampler (halton / sobol): the dimension Get1D / Get2D continue at after the camera sample's five —
    // arrayEndDim (sampler.cpp:141-142, 180-194): 5, or past the sample arrays an integrator requested
    int firstDimension;
};

// Per-path state, structure of arrays, `capacity` entries each.
struct PathState {
    uint32_t *pixel;     // pixel index inside the sample bounds (row-major)
    uint32_t *sample;    // sample index (PCG32 sequence = sample << 32 | pixel)
    uint64_t *rng;       // PCG32 state | halton: sample index (bits 0..47) and next dimension (48..63)
    float4 *rayO;        // xyz origin, w = pFilm.x
    float4 *rayD;        // xyz direction, w = pFilm.y
    float4 *beta;        // xyz throughput, w = etaScale
    float4 *L;           // xyz radiance, w = flags (bounces | specular << 16) as int bits
    int2 *hit;           // x = BVH leaf slot (-1 miss), y = t bits
    // direct light of the last shaded vertex, waiting for its rays (cast together with the next
    // bounce's closest-hit rays): pendS = light-sampled term, zeroed if the shadow ray is occluded;
    // pendM = BSDF-sampled term, written when the MIS ray reaches the sampled light.  The next
    // shade (or the film kernel) folds L += pendS + pendM.
    float4 *pendS;
    float4 *pendM;
    // Integrator "directlighting": the vertices whose SpecularTransmit is still to be traced while
    // the path follows SpecularReflect first (integrator.cpp:341-420 recursion, depth-first per path);
    // kDirectStack entries per path: ray origin + depth, ray direction + leaf slot, throughput
    float4 *stackO, *stackD, *stackBeta;
    // scenes with image textures or bump maps: 8 float4 per path, written by textureKernel for the hit about
    // to be shaded — the IileMaterial with its textures evaluated (96 B), shading normal, shading tangent
    float4 *texRecord;
    // the direct integrators over a scene with textures (null otherwise): rays carry differentials through the
    // specular recursion.  rayDiff: 4 float4 per path, the current ray's rxOrigin (w != 0: the ray has differentials),
    // ryOrigin, rxDirection, ryDirection.  dirRec: kDirRec float4 per path, written by textureDirectKernel for the
    // vertex about to be shaded — the material with its textures evaluated (6), shading normal, shading tangent,
    // dpdx (w != 0: differentials to hand on), dpdy, dndx, dndy, the ray's rxDirection, ryDirection.  stackRec:
    // the dirRec of every parked vertex (kDirectStack per path).
    float4 *rayDiff, *dirRec, *stackRec;
};
constexpr int kDirectStack = 8;  // maxdepth <= 9
constexpr int kDirRec = 14;

struct ShadowItem {  // 32 B; the radiance it gates sits in PathState::pendS[slot]
    float4 o;        // xyz origin, w = slot (int bits)
    float4 d;        // xyz direction, w = tMax
};
struct MisItem {  // 48 B
    float4 o;        // xyz origin, w = slot (int bits)
    float4 d;        // xyz direction, w = light index (int bits)
    float4 contrib;  // xyz: beta * f * weight / pdf / lightPickPdf; multiplied by Le on a match
                     // w: -1, or (int bits) the leaf slot of the sampled light's primitive when the ray reached it
};

// Shading classes: between the closest-hit and the shade kernel the active queue is partitioned
// by what the shade kernel will run for the path (classifyKernel), so that warps — and whole SMs
// at any moment — execute one material's code.
enum : int { kClassMiss = 0, kClassMatte = 1, kClassPlastic = 2, kClassMirror = 3, kClassGlass = 4,
             kClassOther = 5, kNumClasses = 6 };
constexpr int kClassCounter0 = 8;  // counters[8 .. 8+kNumClasses)
constexpr int kNumCounters = 16;

struct Queues {
    uint32_t *active[2];
    uint32_t *byClass[kNumClasses];  // path slots per shading class, filled by classifyKernel
    ShadowItem *shadow;
    MisItem *mis;
    uint32_t *done;
    // directlighting: several shadow / MIS rays per path and vertex, so the radiance a shadow ray
    // gates travels with its queue index (shadowContrib[k]) and is added to PathState::L by the cast
    // itself when the ray is unoccluded; MIS terms are added by misResolveKernel
    float4 *shadowContrib;
    int directMode;
    // counters[0..1] active ping/pong, [2] shadow, [3] mis, [4] done, [8..13] shading classes
    uint32_t *counters;
    unsigned long long *stats;  // [0] closest rays, [1] shadow rays, [2] camera rays
};

struct HemiOutputs {
    const float *cams;       // per hemisphere: cameraToWorld[16], WorldToCamera.mInv[16] (normals)
    float *normals;          // n*size*size*3 (camera-space normal of the first hit)
    float *distance;         // n*size*size
};

// iile_render_dump_rays_dev: where the sampled rays of a render go (device pointers)
struct RayDump {
    IileRay *rays;
    uint8_t *kinds;              // 0 camera ray, 1 bounce ray, 2 MIS ray (closest hit), 3 shadow ray (any hit)
    unsigned long long cap;
    unsigned long long *count;   // rays offered so far (may run past cap)
    uint32_t stride, phase;      // item i of a cast launch is recorded when i % stride == phase % stride
};

struct WavefrontLaunch {
    cudaStream_t stream;
    int blocks;  // persistent grid: multiple of the SM count
};

cudaError_t wfGenerate(const WavefrontLaunch &l, const RenderConsts &rc, PathState st, Queues q,
                       uint64_t firstWork, uint32_t nWork, HemiOutputs hemi);
// one launch for the three ray queues: closest hits of active[cur], then the MIS rays and the
// shadow rays left by the previous shade
cudaError_t wfCast(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q, int cur);
cudaError_t wfDumpRays(const WavefrontLaunch &l, PathState st, Queues q, int cur, RayDump dump);
cudaError_t wfMisResolve(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q);
cudaError_t wfClassify(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q, int cur);
// genericMask / fastMask: shading classes (bit per kClass*) served by the generic kernel / by their own
// per-material kernel; *launches is advanced by the number of kernels enqueued
cudaError_t wfShade(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st,
                    Queues q, int cur, HemiOutputs hemi, int genericMask, int fastMask, bool simpleLights,
                    uint64_t *launches);
// texture stage: between wfClassify and wfShade, for scenes whose materials carry textures
cudaError_t wfTexturesDirect(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st, Queues q,
                             int cur);
cudaError_t wfTextures(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st, Queues q,
                       HemiOutputs hemi);
cudaError_t wfShadeDirect(const WavefrontLaunch &l, const DeviceScene &sc, const RenderConsts &rc, PathState st,
                          Queues q, int cur, int strategy);
cudaError_t wfFilm(const WavefrontLaunch &l, const RenderConsts &rc, PathState st, Queues q, float4 *film,
                   float *hemiIntensity);
// Reorders active[cur] so that neighbouring entries are rays that start in the same cell of a 32^3
// grid over the scene bound and point into the same octant (counting sort on an 18-bit key, queue
// length read on the device).  Only the order of the queue changes.
constexpr int kSortBins = 1 << 18;
cudaError_t wfSortActive(const WavefrontLaunch &l, const DeviceScene &sc, PathState st, Queues q, int cur,
                         uint32_t *keys, uint32_t *hist, uint32_t *sorted);
cudaError_t wfResetCounters(const WavefrontLaunch &l, Queues q, int next);

// spatial light grid (SpatialLightDistribution::ComputeDistribution, lightdistrib.cpp:231-299)
cudaError_t buildSpatialLightGrid(const DeviceScene &sc, const float *halton /*128*5*/, float *table,
                                  cudaStream_t stream);

// IisptRenderRunner's map transforms around the network (hemi_maps.cu; iisptrenderrunner.cpp:1041-1136)
cudaError_t hemiToNN(const float4 *film, const float *normals, const float *distance, size_t n, int size,
                     float filmScale, float *nn, float *means, cudaStream_t stream);
cudaError_t hemiResolve(const float4 *film, size_t nPix, float filmScale, float *rgb, cudaStream_t stream);
cudaError_t hemiFromNN(const float *nnOut, const float *means, size_t n, int size, float *intensity,
                       cudaStream_t stream);

// ---- one task of IisptRenderRunner::run (iispt_task.cu) ----------------------------------------------
struct IisptTaskDev {
    int x0, y0, x1, y1, tilesize, hemiSize;
    int gridW, gridH;              // hemi points of the task
    int stage;                     // 0: work items are the hemi points, 1: the film pixels
    uint32_t samplerCounter, samplerSample;
    uint64_t rngSeed;
};
enum : int { kIisptActive = 0, kIisptNoIntersection = 1, kIisptBlack = 2, kIisptFound = 3 };
struct IisptVertex {               // find_intersection's result for one camera sample
    float4 o, d;                   // the ray that found the vertex (f_ray)
    float4 beta;
    float4 aux;                    // origin of isect.SpawnRay(turned normal): the hemi camera's position
    float4 n;                      // the geometric normal turned against the ray: the hemi camera's axis
    float4 p;                      // isect.p
    uint64_t rngState, rngInc;     // the camera sample's PCG32 stream
    int status, prim;              // kIispt*, scene-order primitive of the vertex
    int pad[2];
};
struct IisptScratch {
    IisptVertex *verts;
    IileRay *rays[2];
    uint32_t *active[2];
    IileHit *hits;
    uint32_t *counter;
    float4 *texRecord;             // 8 per item, scenes with textures only (else null)
};
void iisptGrid(const IisptTaskDev &t, int *gridW, int *gridH);
cudaError_t iisptFindVertices(const DeviceScene &sc, const RenderConsts &rc, const IisptTaskDev &t, uint32_t n,
                              IisptScratch scratch, cudaStream_t stream, uint64_t *launches);
cudaError_t iisptHemiPoints(uint32_t n, const IisptVertex *verts, float *pos, float *dir, uint8_t *valid,
                            cudaStream_t stream);
cudaError_t iisptGather(const DeviceScene &sc, const RenderConsts &rc, const IisptTaskDev &t, uint32_t n,
                        const IisptVertex *verts, const float *hemiPos, const float *hemiDir, const int *hemiFilm,
                        const float *cams, const float *films, float4 *out, float4 *texRecord, cudaStream_t stream);

}  // namespace iile
