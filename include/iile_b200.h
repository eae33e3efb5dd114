/* iile_b200.h — C ABI of the B200-native rendering hot path of pbrt-v3-IILE.
 *
 * The reference has no FFI: its "plugins" are C++ classes picked by name in
 * src/core/api.cpp and driven from pbrtWorldEnd() (api.cpp:1632-1692).  This
 * header is the seam a maintainer would bind instead of
 *   - RenderOptions::MakeScene  -> BVHAccel ctor        (api.cpp:1694-1703, bvh.cpp:183-225)
 *   - Aggregate::Intersect / IntersectP                 (primitive.h:118-126, bvh.cpp:662-738)
 *   - Integrator::Render (SamplerIntegrator / PathIntegrator) (integrator.h:53-58,
 *     integrator.cpp:227-339, path.cpp:64-194)
 *   - Film::MergeFilmTile / WriteImage                  (film.cpp:135-148,187-235)
 *   - IISPTdIntegrator::RenderView                      (iispt_d.cpp:345-461)
 * Every entry point is blocking, returns 0 on success or a negative IILE_E_*
 * code, and never throws or aborts across the boundary (iile_last_error()
 * returns the message of the calling thread's last failure).  Plain pointers
 * and sizes only; "_dev" variants take CUDA device pointers plus a
 * cudaStream_t passed as void*.  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with IILE_E_CUDA.
 */
#ifndef IILE_B200_H
#define IILE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IILE_ABI_VERSION 4

enum {
    IILE_OK = 0,
    IILE_E_INVALID = -1, /* bad argument / malformed description */
    IILE_E_PARSE = -2,   /* .pbrt front-end error (unsupported directive, syntax) */
    IILE_E_CUDA = -3,    /* CUDA runtime failure or no device */
    IILE_E_IO = -4,
    IILE_E_UNSUPPORTED = -5
};

/* ---- ray casting records (Ray: geometry.h:869-890) ----------------------- */
typedef struct IileRay {
    float o[3];
    float tMax;
    float d[3];
    float time;
} IileRay; /* 32 B */

/* prim = index of the primitive in scene-description order, -1 on a miss.
 * t = ray.tMax after BVHAccel::Intersect; b1,b2 = barycentrics of the hit
 * (triangle.cpp:266-270; b0 = 1-b1-b2 only approximately — shading recomputes
 * all three).  For spheres b1,b2 hold the surface (u,v). */
typedef struct IileHit {
    int32_t prim;
    float t;
    float b1;
    float b2;
} IileHit; /* 16 B */

/* LinearBVHNode, bvh.cpp:95-104: depth-first order, first child = index+1. */
typedef struct IileBVHNode {
    float bmin[3];
    float bmax[3];
    int32_t offset;   /* leaf: first primitive (in BVH order); interior: second child */
    uint16_t nPrims;  /* 0 => interior */
    uint8_t axis;
    uint8_t pad;
} IileBVHNode; /* 32 B */

/* ---- flat scene description ---------------------------------------------- */
enum { IILE_PRIM_TRIANGLE = 0, IILE_PRIM_SPHERE = 1 };
enum {
    IILE_PRIM_HAS_N = 1u,        /* mesh carries per-vertex normals */
    IILE_PRIM_HAS_UV = 2u,       /* mesh carries per-vertex uv */
    IILE_PRIM_FLIP_NORMAL = 4u   /* reverseOrientation ^ transformSwapsHandedness */
};
typedef struct IilePrim {
    int32_t type;      /* IILE_PRIM_* */
    int32_t v[3];      /* triangle: vertex indices into P/N/UV; sphere: v[0] = sphere index */
    int32_t material;  /* index into materials */
    int32_t light;     /* index into lights of this primitive's area light, -1 if none */
    uint32_t flags;
    uint32_t pad;
} IilePrim; /* 32 B */

/* Shape "sphere" (sphere.cpp:317-327) and Shape "disk" (disk.cpp:163-173), kept in object space like
 * the reference.  flags bit2 marks a disk, whose height lives in zMin and inner radius in zMax; bit3 a cylinder
 * (cylinder.cpp:222-233: radius, zMin, zMax, phiMax as named). */
typedef struct IileSphere {
    float objectToWorld[16]; /* row-major 4x4 */
    float worldToObject[16];
    float radius, zMin, zMax, thetaMin, thetaMax, phiMax; /* as Sphere's ctor derives them */
    uint32_t flags; /* bit0 reverseOrientation, bit1 transformSwapsHandedness, bit2 disk, bit3 cylinder */
    float phiMaxDegrees; /* the "phimax" parameter as given (degrees) */
} IileSphere;

enum { IILE_MAT_MATTE = 0, IILE_MAT_PLASTIC = 1, IILE_MAT_MIRROR = 2, IILE_MAT_GLASS = 3, IILE_MAT_UBER = 4,
       IILE_MAT_SUBSTRATE = 5 /* substrate.cpp: Kd, Ks, u/vRoughness */,
       IILE_MAT_TRANSLUCENT = 6 /* translucent.cpp: Kd, Ks, roughness; reflect in Kr, transmit in Kt */ };
/* Constant-texture materials (matte.cpp:45-71, plastic.cpp:45-83, mirror.cpp:45-64,
 * glass.cpp:45-110 smooth and rough, uber.cpp:45-137, substrate.cpp:45-83, translucent.cpp:45-105).  Unused fields are zero.  ABI 2 appended
 * Kr / opacity (96 B; ABI 1 and IILEFLT1 files carried the first 64 B). */
typedef struct IileMaterial {
    int32_t type;
    int32_t remapRoughness;
    float Kd[3];
    float Ks[3]; /* plastic Ks; mirror/glass Kr */
    float Kt[3]; /* glass */
    float sigma;      /* matte */
    float roughness;  /* plastic */
    float eta;        /* glass */
    float uRoughness, vRoughness; /* glass; uber: roughness resolved per axis (uber.cpp:71-79) */
    float Kr[3];      /* uber specular reflection */
    float opacity[3]; /* uber */
    int32_t texSet;   /* ABI 4: 0 = every parameter is the constant above; k > 0: entry k-1 of
                         IileSceneDesc::materialTextures names the textures that replace them */
    float pad;
} IileMaterial; /* 96 B */

/* ---- textures (ABI 4): Texture "..." "float|spectrum" "constant" / "scale" / "imagemap" ------------- */
enum { IILE_TEX_CONSTANT = 0, IILE_TEX_IMAGEMAP = 1, IILE_TEX_SCALE = 2,
       IILE_TEX_MIX = 3 /* MixTexture, mix.h:49-66: (1 - amount) * tex1 + amount * tex2 */,
       IILE_TEX_CHECKERBOARD = 4 /* Checkerboard2DTexture over tex1 / tex2 with UVMapping2D (checkerboard.h:52-107);
                                    `trilinear` != 0 = "aamode" "none", else the closed-form box filter */ };
#define IILE_TEX_MAX_DEPTH 3 /* operators (scale / mix / checkerboard) nested at most this deep over constant / imagemap leaves */
enum { IILE_WRAP_REPEAT = 0, IILE_WRAP_BLACK = 1, IILE_WRAP_CLAMP = 2 }; /* ImageWrap, mipmap.h:52 */
/* ConstantTexture (constant.h), ScaleTexture (scale.h:48-64) and MixTexture (mix.h:49-66) over operands that come
 * EARLIER in the texture array (so the records form a DAG; at most IILE_TEX_MAX_DEPTH operators deep),
 * ImageTexture with UVMapping2D (imagemap.h:76-120, texture.cpp UVMapping2D::Map). */
typedef struct IileTexture {
    int32_t type;          /* IILE_TEX_* */
    int32_t channels;      /* 1: Texture<Float>, 3: Texture<Spectrum> */
    float value[3];        /* constant */
    int32_t tex1, tex2;    /* scale / mix: operand textures (same channel count, smaller indices) */
    int32_t image;         /* imagemap: index into images; mix: the float texture giving "amount", -1 = value[0] */
    float su, sv, du, dv;  /* imagemap, checkerboard: "uscale" "vscale" "udelta" "vdelta" */
    int32_t wrap;          /* IILE_WRAP_* */
    int32_t trilinear;     /* "trilinear" (false = EWA, mipmap.h:343-369) */
    float maxAnisotropy;
    float pad;
} IileTexture; /* 64 B */
/* A MIPMap pyramid (mipmap.h:147-237) as the reference builds it: the image flipped in y, converted
 * (scale, inverse gamma; luminance for a float texture: imagemap.h:98-106), resampled to powers of two
 * with the Lanczos weights, then box-filtered level by level.  Level l holds max(1, width >> l) x
 * max(1, height >> l) texels of `channels` floats, row-major, level after level from texelOffset on. */
typedef struct IileImage {
    int32_t width, height; /* level 0 (powers of two) */
    int32_t nLevels;
    int32_t channels;      /* 1 or 3 */
    uint64_t texelOffset;  /* index of level 0's first float in the texel pool */
    uint64_t pad;
} IileImage; /* 32 B */
/* Which texture replaces which parameter of an IileMaterial (-1 = the constant stays).  bump = the
 * "bumpmap" float texture of Material::Bump (material.cpp:46-83). */
enum { IILE_TEXSLOT_KD = 0, IILE_TEXSLOT_KS, IILE_TEXSLOT_KT, IILE_TEXSLOT_KR, IILE_TEXSLOT_OPACITY, IILE_TEXSLOT_SIGMA,
       IILE_TEXSLOT_ROUGHNESS, IILE_TEXSLOT_UROUGHNESS, IILE_TEXSLOT_VROUGHNESS, IILE_TEXSLOT_ETA, IILE_TEXSLOT_BUMP,
       IILE_TEXSLOT_COUNT = 12 };
typedef struct IileMaterialTextures {
    int32_t slot[IILE_TEXSLOT_COUNT];
} IileMaterialTextures; /* 48 B */

enum { IILE_LIGHT_POINT = 0, IILE_LIGHT_DIFFUSE_AREA = 1, IILE_LIGHT_SPOT = 2, IILE_LIGHT_DISTANT = 3,
       IILE_LIGHT_INFINITE = 4 /* InfiniteAreaLight with a constant L (no "mapname"), infinite.cpp */ };
/* point.cpp, diffuse.cpp, spot.cpp, distant.cpp, infinite.cpp.  ABI 2 appended everything after `area`
 * (ABI 1 and IILEFLT1 files carried the first 40 B). */
typedef struct IileLight {
    int32_t type;
    int32_t prim;      /* area: primitive (scene order) that emits */
    float pos[3];      /* point / spot: pLight = LightToWorld(0,0,0) (point.cpp:47, spot.cpp:48);
                          distant: wLight = Normalize(LightToWorld(dir)) (distant.cpp:46-49) */
    float L[3];        /* point / spot: I*scale; area / distant: L*scale */
    int32_t twoSided;  /* area: bit 0 = "twosided"; bits 8..23 = "samples" - 1 (diffuse.cpp:140-141; ABI 4: files and
                          callers that wrote 0 / 1 here mean one sample).  IILE_LIGHT_TWOSIDED / IILE_LIGHT_NSAMPLES */
    float area;        /* area: Shape::Area() */
    float lightToWorld[16]; /* spot, distant: row-major LightToWorld.m (spot.cpp:117-121) */
    float worldToLight[16]; /* spot: LightToWorld.mInv, what Falloff applies to directions */
    float coneAngle, falloffStart; /* spot: totalWidth and falloffStart in degrees, as the ctor takes them */
    float cosTotalWidth, cosFalloffStart; /* spot: std::cos(Radians(.)) of the two */
    float dir[3];      /* distant: from - to as given (distant.cpp:100) */
    float worldRadius; /* distant, infinite: Bounds3::BoundingSphere of the scene bound, filled at scene creation */
    /* infinite: the Distribution2D InfiniteAreaLight's ctor builds over its 2x2 luminance image of
     * the 1x1 map (infinite.cpp:66-84): conditional func[2][2], cdf[2][3], funcInt[2]; marginal
     * func[2], cdf[3], funcInt */
    float envFunc[2][2], envCdf[2][3], envFuncInt[2], envMargFunc[2], envMargCdf[3], envMargFuncInt;
} IileLight; /* 272 B */

#define IILE_LIGHT_TWOSIDED(l) (((l).twoSided & 1) != 0)
#define IILE_LIGHT_NSAMPLES(l) (1 + (((l).twoSided >> 8) & 0xffff))

enum { IILE_FILTER_BOX = 0, IILE_FILTER_GAUSSIAN = 1, IILE_FILTER_MITCHELL = 2,
       IILE_FILTER_SINC = 3, IILE_FILTER_TRIANGLE = 4,
       /* not a PixelFilter: IisptFilmMonitor::add_n_samples (iisptfilmmonitor.cpp:47-75) as
        * DirectProgressiveIntegrator::RenderOnePass feeds it (directprogressiveintegrator.cpp:137-148) — a sample is
        * added, with weight 1, to the pixel it was drawn FOR, wherever pixel + u rounds to; filterRadius is 0.5 */
       IILE_FILTER_MONITOR = 5 };
/* Sampler "random" (samplers/random.cpp) | "halton" (samplers/halton.cpp:63-122) */
enum { IILE_SAMPLER_RANDOM = 0, IILE_SAMPLER_HALTON = 1, IILE_SAMPLER_SOBOL = 2 };
enum { IILE_CAMERA_PERSPECTIVE = 0, IILE_CAMERA_ORTHOGRAPHIC = 1, IILE_CAMERA_ENVIRONMENT = 2 };
enum { IILE_INTEGRATOR_PATH = 0, IILE_INTEGRATOR_DIRECT = 1,
       IILE_INTEGRATOR_WHITTED = 2 /* WhittedIntegrator, whitted.cpp:44-88: same recursion, unweighted light loop */ };
enum { IILE_DIRECT_ALL = 0, IILE_DIRECT_ONE = 1 };
enum { IILE_LIGHTSTRATEGY_UNIFORM = 0, IILE_LIGHTSTRATEGY_POWER = 1, IILE_LIGHTSTRATEGY_SPATIAL = 2 };
enum { IILE_SPLIT_SAH = 0, IILE_SPLIT_HLBVH = 1, IILE_SPLIT_MIDDLE = 2, IILE_SPLIT_EQUAL = 3 };

/* Everything pbrtWorldEnd's MakeIntegrator() reads from the option block
 * (Camera/Film/PixelFilter/Sampler/Integrator/Accelerator directives). */
typedef struct IileRenderSetup {
    /* Camera "perspective" (perspective.cpp:293-329) | "orthographic" | "environment": cameraType below */
    float cameraToWorld[16];
    float worldToCamera[16];
    float fov, lensRadius, focalDistance;
    float screenWindow[4]; /* xmin xmax ymin ymax */
    float shutterOpen, shutterClose;
    /* Film "image" (film.cpp:259-303) */
    int32_t xRes, yRes;
    float cropWindow[4]; /* x0 x1 y0 y1 */
    float filmScale, maxSampleLuminance, diagonal;
    /* PixelFilter */
    int32_t filterType;
    float filterRadius[2];
    float filterParam[2]; /* gaussian alpha | mitchell B,C | sinc tau */
    /* Sampler "random" | "halton" (samplerType below) */
    int32_t pixelSamples;
    /* Integrator "path" (path.cpp:214-231) */
    int32_t maxDepth;
    float rrThreshold;
    int32_t hasPixelBounds;
    int32_t pixelBounds[4]; /* x0 x1 y0 y1 */
    int32_t lightSampleStrategy;
    /* Accelerator "bvh" (bvh.cpp:740-760) */
    int32_t maxNodePrims;
    int32_t splitMethod;
    char filename[236];
    /* Camera "perspective" (0) | "orthographic" (1, orthographic.cpp) | "environment" (2, environment.cpp);
     * zero in older flat files, like the words below */
    int32_t cameraType;
    /* Integrator "path" (0) | "directlighting" (1: DirectLightingIntegrator, directlighting.cpp:45-99,
     * the Li of IILE's DirectProgressiveIntegrator, directprogressiveintegrator.cpp:22-62) and its
     * "strategy" (0 = "all": UniformSampleAllLights, 1 = "one": UniformSampleOneLight).  Like the two
     * words below these were the tail of filename[256] in older flat files, where they read as zero. */
    int32_t integratorType;
    int32_t directStrategy;
    /* IILE_SAMPLER_*; halton only: "bool samplepixelcenter" (halton.cpp:124-130).  These two
     * words were the tail of filename[256] in flat files written before they existed, where
     * they read as zero = random sampler. */
    int32_t samplerType;
    int32_t samplePixelCenter;
} IileRenderSetup;

typedef struct IileSceneDesc {
    uint32_t nVerts;
    const float *P;   /* 3*nVerts, world space */
    const float *N;   /* 3*nVerts or NULL */
    const float *UV;  /* 2*nVerts or NULL */
    uint32_t nPrims;
    const IilePrim *prims;
    uint32_t nSpheres;
    const IileSphere *spheres;
    uint32_t nMaterials;
    const IileMaterial *materials;
    uint32_t nLights;
    const IileLight *lights;
    /* optional prebuilt accelerator; NULL => built here (SAH, bvh.cpp:236-402) */
    uint32_t nNodes;
    const IileBVHNode *nodes;
    const uint32_t *primOrder; /* nPrims: BVH-order slot -> scene-order primitive */
    IileRenderSetup setup;
    /* ABI 4: textures (all optional) */
    uint32_t nTextures;
    const IileTexture *textures;
    uint32_t nImages;
    const IileImage *images;
    uint64_t nTexels;              /* floats in the texel pool */
    const float *texels;
    uint32_t nMaterialTextures;
    const IileMaterialTextures *materialTextures;
} IileSceneDesc;

typedef struct IileScene IileScene;

typedef struct IileSceneInfo {
    uint32_t nVerts, nPrims, nTriangles, nSpheres, nMaterials, nLights, nNodes;
    float worldMin[3], worldMax[3];
    IileRenderSetup setup;
    int32_t sampleBounds[4];   /* x0 y0 x1 y1, Film::GetSampleBounds (film.cpp:77-83) */
    int32_t croppedBounds[4];  /* x0 y0 x1 y1, film.cpp:55-59 */
    double buildSeconds;       /* host BVH build */
} IileSceneInfo;

const char *iile_last_error(void);
int iile_abi_version(void);
/* number of CUDA devices visible (0 => compute entry points fail) */
int iile_device_count(void);
int iile_set_device(int device);

/* ---- scene: parser + flattening + BVH + upload ---------------------------- */
/* ParseFile + pbrtWorldEnd's MakeScene for the supported directive subset
 * (parser.cpp:712, api.cpp:1632-1703).  upload=0 keeps the scene host-only
 * (no CUDA needed: used to inspect / save the flattened scene). */
int iile_scene_load_pbrt(const char *path, int upload, IileScene **out);
int iile_scene_create(const IileSceneDesc *desc, int upload, IileScene **out);
/* flat-scene file = the IileSceneDesc arrays serialised */
int iile_scene_save_flat(const IileScene *scene, const char *path);
int iile_scene_load_flat(const char *path, int upload, IileScene **out);
void iile_scene_destroy(IileScene *scene);
int iile_scene_info(const IileScene *scene, IileSceneInfo *out);
/* copy out the accelerator (for checking against the reference's own tree) */
int iile_scene_get_nodes(const IileScene *scene, IileBVHNode *nodes, uint32_t cap);
int iile_scene_get_prim_order(const IileScene *scene, uint32_t *order, uint32_t cap);
/* override the integrator/sampler/film options read from the file */
int iile_scene_set_setup(IileScene *scene, const IileRenderSetup *setup);

/* ---- BVHAccel::Intersect / IntersectP -------------------------------------- */
/* host buffers: copies in, casts, copies out */
int iile_intersect_closest(IileScene *scene, const IileRay *rays, size_t n, IileHit *hits);
int iile_intersect_any(IileScene *scene, const IileRay *rays, size_t n, uint8_t *occluded);
/* device buffers, asynchronous on `stream` */
int iile_intersect_closest_dev(IileScene *scene, const IileRay *rays, size_t n, IileHit *hits,
                               void *stream);
int iile_intersect_any_dev(IileScene *scene, const IileRay *rays, size_t n, uint8_t *occluded,
                           void *stream);
/* cfg-2 workload generator: chunk c of 2^14 rays draws from
 * PCG32 sequence c, five UniformFloat per ray (ox oy oz u0 u1); origin inside
 * [bmin,bmax], direction UniformSampleSphere(u0,u1), tMax = inf. */
int iile_generate_rays_dev(const float bmin[3], const float bmax[3], uint64_t firstRay, size_t n,
                           IileRay *rays, void *stream);

/* ---- SamplerIntegrator::Render / PathIntegrator::Li ------------------------ */
typedef struct IileRenderParams {
    int32_t sampleBegin, sampleEnd; /* samples [begin,end) of every pixel; end<=0 => pixelSamples */
    int32_t tileBegin, tileEnd;     /* 16x16 tiles in Render's order; end<=0 => all (rank partition) */
    int32_t clearFilm;              /* zero the film before accumulating */
    int32_t flags;                  /* IILE_RENDER_* */
    uint64_t maxPathsInFlight;      /* 0 => default */
} IileRenderParams;
/* with a non-NULL `stats`: CUDA events between the launches of the render, summed per kernel family
 * into IileRenderStats.castSeconds / shadeSeconds / otherSeconds (measurement only) */
enum { IILE_RENDER_TIME_KERNELS = 1 };

typedef struct IileRenderStats {
    uint64_t cameraRays;     /* samples traced                        */
    uint64_t closestRays;    /* Scene::Intersect calls (incl. MIS rays) */
    uint64_t shadowRays;     /* Scene::IntersectP calls                */
    uint64_t launches;       /* kernels launched                       */
    double deviceSeconds;    /* CUDA-event time of the kernels (host variant only) */
    /* IILE_RENDER_TIME_KERNELS: device time of the traversal launches (BVHAccel::Intersect / IntersectP),
     * of the shading launches (Li's body; includes the class partition in front of it) and of the rest
     * (camera rays, MIS resolve, film, counter resets); zero otherwise */
    double castSeconds, shadeSeconds, otherSeconds;
} IileRenderStats;

/* film = {R,G,B,filterWeightSum} float4 per pixel of the cropped pixel bounds,
 * row-major: FilmTilePixel sums (film.h:153-193) before the RGB->XYZ merge. */
int iile_render_dev(IileScene *scene, const IileRenderParams *params, float *filmRGBW,
                    IileRenderStats *stats, void *stream);
int iile_render(IileScene *scene, const IileRenderParams *params, float *filmRGBW,
                IileRenderStats *stats);
/* Measurement aid: iile_render_dev that also copies every `stride`-th ray its traversal launches trace
 * into rays[cap] (device), kinds[cap] (device; 0 camera ray, 1 bounce ray, 2 BSDF-sampled MIS ray —
 * closest hit —, 3 shadow ray — any hit).  *count (host) = rays offered; min(*count, cap) were stored.
 * The rays are what Scene::Intersect / IntersectP see inside PathIntegrator::Li (path.cpp:88,
 * integrator.cpp:160-211): cast them stand-alone (iile_intersect_*_dev) or hand them to the oracle to
 * count the reference traversal's node and triangle fetches for a rendered config.  Synchronises. */
int iile_render_dump_rays_dev(IileScene *scene, const IileRenderParams *params, float *filmRGBW,
                              IileRay *rays, uint8_t *kinds, uint64_t cap, uint32_t stride,
                              uint64_t *count, void *stream);
/* Film::MergeFilmTile + WriteImage arithmetic on the host (film.cpp:135-148,187-226):
 * RGB->XYZ, XYZ->RGB, /weight, clamp at 0, *scale.  rgb = 3 floats per pixel. */
int iile_film_resolve(const IileScene *scene, const float *filmRGBW, float *rgb);
/* WriteImage by extension: .pfm (imageio.cpp:437-482) or .exr (half RGBA, imageio.cpp:164-189) */
int iile_write_image(const char *path, const float *rgb, int xRes, int yRes, int fullX, int fullY,
                     int offX, int offY);

/* ---- device buffers for hosts without a CUDA toolchain (cgo / JNI / plain C++) ----------- */
/* Memory on the calling thread's current device (iile_set_device), for the "_dev" entry points. */
int iile_dev_malloc(size_t bytes, void **ptr);
int iile_dev_free(void *ptr);
int iile_dev_upload(void *dst, const void *src, size_t bytes);   /* host -> device, blocking */
int iile_dev_download(void *dst, const void *src, size_t bytes); /* device -> host, blocking */
int iile_dev_synchronize(void *stream);

/* ---- Film::MergeFilmTile across GPUs (film.cpp:135-148) -------------------- */
/* One process (or host thread) per GPU renders a sample or tile range of the same image into
 * its own {R,G,B,w} film (IileRenderParams); the films are then summed with one NCCL collective
 * over NVLink.  `comm` is a plain ncclComm_t: either one the host already owns, or one made here
 * (rank 0 calls iile_comm_unique_id, ships the 128 bytes to the other ranks by any channel, every
 * rank calls iile_comm_create on its own device).  NCCL is bound at run time (libnccl.so.2). */
#define IILE_COMM_ID_BYTES 128
int iile_comm_nccl_version(int *version);
int iile_comm_unique_id(uint8_t id[IILE_COMM_ID_BYTES]);
int iile_comm_create(const uint8_t id[IILE_COMM_ID_BYTES], int nRanks, int rank, void **comm);
int iile_comm_destroy(void *comm);
/* In-place sum of nPixels float4 film pixels on `stream` (asynchronous): onto rank `root`, or
 * onto every rank when root < 0.  comm == NULL or a one-rank communicator: no-op. */
int iile_film_reduce_dev(float *filmRGBW, size_t nPixels, int root, void *comm, void *stream);

/* ---- IISPTdIntegrator::RenderView over a batch of hemispheres -------------- */
/* pos/dir: 3 floats per hemisphere (camera position and axis, hemispheric.cpp:108-158).
 * intensity: n*size*size*3, normals: n*size*size*3, distance: n*size*size (host). */
int iile_render_hemispheres(IileScene *scene, size_t n, const float *pos, const float *dir,
                            int size, int spp, float *intensity, float *normals,
                            float *distance, IileRenderStats *stats);

/* The same batch with IisptRenderRunner's map transforms around the network
 * (iisptrenderrunner.cpp:1041-1136) done on the device, replacing the per-hemisphere stdio round
 * trip of IisptNnConnector::communicate (iisptnnconnector.cpp:125-148):
 *   to_nn   = RenderView + normalizeMapsDownstream -> nnInput [n][7][size][size] (device; planes
 *             R,G,B intensity, x,y,z normal, distance; rows in ImageFilm order = camera row
 *             size-1-Y, what main_stdio_net.read_input builds) and means [n][3] (device)
 *   from_nn = transformMapsUpstream: nnOutput [n][3][size][size] + means -> intensity
 *             [n][size][size][3] (device, ImageFilm order)
 * pos/dir are host pointers.  Asynchronous on `stream`; a non-NULL `stats` synchronises once. */
int iile_hemispheres_to_nn_dev(IileScene *scene, size_t n, const float *pos, const float *dir,
                               int size, int spp, float *nnInput, float *means,
                               IileRenderStats *stats, void *stream);
int iile_hemispheres_from_nn_dev(size_t n, int size, const float *nnOutput, const float *means,
                                 float *intensity, void *stream);

/* ---- one task of IisptRenderRunner::run (iisptrenderrunner.cpp:216-596) ----------------------------- */
/* The pixel rectangle [x0,x1) x [y0,y1) of an IisptScheduleMonitorTask with its grid of hemi points every
 * `tilesize` pixels (and always on the last column / row).  Camera samples are drawn from one PCG32 stream
 * each: sequence = samplerSample << 32 | c, c = samplerCounter + 1, + 2, ... in the order run() calls
 * sampler_next_pixel() (hemi points row by row, then the film pixels row by row).  The gather's
 * IisptRng stream of film pixel j (row-major inside the task) is RNG(rngSeed + j). */
typedef struct IileIisptTask {
    int32_t x0, y0, x1, y1;
    int32_t tilesize;
    int32_t hemiSize;        /* PbrtOptions.iisptHemiSize */
    uint32_t samplerCounter; /* IisptRenderRunner::sampler_pixel_counter.x before the task */
    uint32_t samplerSample;
    uint64_t rngSeed;
} IileIisptTask;
/* number of hemi points of a task: gridW * gridH */
int iile_iispt_grid(const IileIisptTask *task, int32_t *gridW, int32_t *gridH);
/* Stage 1 (run :252-325): find_intersection for the main camera's ray through every hemi point ->
 * position and axis of its HemisphericCamera (pos / dir: 3 floats per hemi point, host) and valid[h] = 0
 * where the reference stores a null camera (no intersection, or beta.y() <= 0).  The valid ones, in
 * order, are what iile_hemispheres_to_nn_dev renders. */
int iile_iispt_hemi_points(IileScene *scene, const IileIisptTask *task, float *pos, float *dir, uint8_t *valid);
/* Stage 3 (run :404-570): per film pixel find_intersection, compute_fpixel_weights, sample_hemisphere /
 * estimate_direct against the network's films.  pos / dir / valid: stage 1's output (host).  nnIntensity
 * (device): [number of valid hemi points][hemiSize][hemiSize][3], the output of
 * iile_hemispheres_from_nn_dev for the valid hemi points in order.  out (device): one float4 per film
 * pixel of the task, row-major: {f_beta * L, sample weight} — what run() hands to
 * IisptFilmMonitor::add_n_samples (weight 0.5; 0 where the pixel adds no sample). */
int iile_iispt_gather_dev(IileScene *scene, const IileIisptTask *task, const float *pos, const float *dir,
                          const uint8_t *valid, const float *nnIntensity, float *out, void *stream);

#ifdef __cplusplus
}
static_assert(sizeof(IileRay) == 32 && sizeof(IileHit) == 16 && sizeof(IileBVHNode) == 32, "ABI");
static_assert(sizeof(IileTexture) == 64 && sizeof(IileImage) == 32 && sizeof(IileMaterialTextures) == 48, "ABI");
static_assert(sizeof(IilePrim) == 32 && sizeof(IileSphere) == 160 && sizeof(IileMaterial) == 96 &&
              sizeof(IileLight) == 272 && sizeof(IileRenderSetup) == 520, "ABI");
#endif
#endif /* IILE_B200_H */
