// C ABI: render entry points.  Host-side driver of the wavefront kernels in
// wavefront.cu — the part of SamplerIntegrator::Render (integrator.cpp:227-339)
// that is not arithmetic: camera / film constants (camera.h:95-110,
// perspective.cpp:50-75, film.cpp:45-83), the light-sampling tables of
// CreateLightSampleDistribution (lightdistrib.cpp:47-81, 95-122, 231-299), the
// work decomposition into batches of paths, and the launch sequence of one
// bounce.  A whole render is enqueued on one stream without host
// synchronisation: every kernel reads its queue length from device memory, and
// the bounce loop runs maxDepth+1 times, which is the longest a path can live
// (path.cpp:104).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "scene_handle.h"
#include "wavefront.h"

using namespace iile;

// Ray binning before the deeper bounces' casts (wfSortActive): measured on B200 and left OFF — it
// costs 30 % (profiles/r5a_ab_ray_sort.txt: cfg1 411 -> 280, cfg3 201 -> 147 Msamples/s).  The active
// queue's order is (nearly) the order of the path-state slots, which is what keeps the 88 B/path state
// reads and writes of the cast, classify and shade kernels coalesced; a queue sorted by ray origin
// turns every one of them into a gather.  Binning would have to move the state, not the indices.
#ifndef IILE_RAY_SORT
#define IILE_RAY_SORT 0
#endif

namespace {

// wavefront steps a batch may add for paths that crossed interface-only primitives (each crossing is a step without
// a vertex); the reference has no bound, a path that crosses more often than this is filmed with what it has
constexpr int kMaxInterfaceSteps = 256;

struct RenderState {
    size_t capacity = 0;
    PathState st{};
    Queues q{};
    std::vector<void *> allocs;
    float *lightTable = nullptr;
    float *hemiCams = nullptr;
    size_t hemiCamsCap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int smCount = 148;
    uint64_t lastLaunches = 0;  // kernels enqueued by the most recent iile_render_dev
    int tableStrategy = -1;     // light-sampling strategy the device table was built for
    HaltonTables *haltonTables = nullptr;  // Sampler "halton": primes + digit permutations
    uint16_t *haltonPerms = nullptr;
    uint32_t *sobolMatrices = nullptr;   // Sampler "sobol": SobolMatrices32, uploaded once
    uint64_t *sobolVdc = nullptr;        // the two VdC rows of the current resolution
    int sobolVdcLog2 = -1;
    // ray binning before the cast of the deeper bounces (wfSortActive)
    uint32_t *sortKeys = nullptr, *sortHist = nullptr, *sortTmp = nullptr;
    // Integrator "directlighting": per-path stacks and the larger ray queues (several lights per vertex)
    float4 *stackO = nullptr, *stackD = nullptr, *stackBeta = nullptr, *shadowContrib = nullptr;
    float4 *rayDiff = nullptr, *dirRec = nullptr, *stackRec = nullptr;  // direct integrators over textured scenes
    ShadowItem *directShadow = nullptr;
    MisItem *directMis = nullptr;
    size_t directPaths = 0, directItems = 0;
    // host-buffer render (iile_render): persistent device film + pinned double buffer for the read-back
    float *hostFilm = nullptr;
    size_t hostFilmBytes = 0;
    void *pinned[2] = {nullptr, nullptr};
    cudaEvent_t pinnedDone[2] = {nullptr, nullptr};
    cudaStream_t copyStream = nullptr;
    // IILE_RENDER_TIME_KERNELS: one event in front of every launch (pool reused across renders)
    std::vector<cudaEvent_t> timingEvents;
    // iile_render_dump_rays_dev: set for the duration of that call
    RayDump dump{};
    // Batch lanes: a render of several batches runs them round-robin on kMaxLanes internal streams, each
    // with its own path state and queues (aux[k - 1] for lane k; lane 0 is this object), so that the
    // under-filled tail of one batch's launches overlaps the bulk of the next batch's.
    RenderState *aux[3] = {nullptr, nullptr, nullptr};
    cudaStream_t laneStream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t laneDone[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t forkEvent = nullptr;
    int lanesUsed = 1;  // by the most recent render: readStats sums the lanes' ray counters
};
constexpr int kMaxLanes = 4;

constexpr size_t kReadbackChunk = (size_t)8 << 20;

// Device -> pageable host memory through two pinned chunks: the DMA of chunk k overlaps the CPU
// copy of chunk k-1 out of the staging buffer (a plain cudaMemcpy into pageable memory does the two
// one after the other; for the 133 MB cfg5 film that was a fifth of the end-to-end step).
cudaError_t downloadPipelined(RenderState *rs, void *dst, const void *src, size_t bytes, cudaEvent_t after) {
    cudaError_t e = cudaSuccess;
    {   // a page-locked destination (cudaHostAlloc / cudaHostRegister) takes the DMA directly
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
            e = cudaEventSynchronize(after);
            return e == cudaSuccess ? cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) : e;
        }
        cudaGetLastError();  // unregistered host memory reports an error on older drivers: not ours
    }
    if (!rs->copyStream) {
        e = cudaStreamCreateWithFlags(&rs->copyStream, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaMallocHost(&rs->pinned[b], kReadbackChunk);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rs->pinnedDone[b], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) return e;
    }
    e = cudaStreamWaitEvent(rs->copyStream, after, 0);
    const size_t nChunks = (bytes + kReadbackChunk - 1) / kReadbackChunk;
    auto span = [&](size_t k) { return std::min(kReadbackChunk, bytes - k * kReadbackChunk); };
    for (size_t k = 0; k <= nChunks && e == cudaSuccess; ++k) {
        if (k < nChunks) {
            const int b = (int)(k & 1);
            e = cudaMemcpyAsync(rs->pinned[b], (const char *)src + k * kReadbackChunk, span(k), cudaMemcpyDeviceToHost,
                                rs->copyStream);
            if (e == cudaSuccess) e = cudaEventRecord(rs->pinnedDone[b], rs->copyStream);
        }
        if (k >= 1 && e == cudaSuccess) {
            const int b = (int)((k - 1) & 1);
            e = cudaEventSynchronize(rs->pinnedDone[b]);
            if (e == cudaSuccess) std::memcpy((char *)dst + (k - 1) * kReadbackChunk, rs->pinned[b], span(k - 1));
        }
    }
    return e;
}

RenderState *stateOf(IileScene *s) { return static_cast<RenderState *>(s->renderState); }

template <typename T>
cudaError_t devAlloc(RenderState *rs, T **p, size_t n) {
    *p = nullptr;
    cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) rs->allocs.push_back(*p);
    return e;
}

void freePathBuffers(RenderState *rs) {
    for (void *p : rs->allocs) cudaFree(p);
    rs->allocs.clear();
    rs->capacity = 0;
}

void freeDirectBuffers(RenderState *rs) {
    for (void *p : {(void *)rs->stackO, (void *)rs->stackD, (void *)rs->stackBeta, (void *)rs->shadowContrib,
                    (void *)rs->directShadow, (void *)rs->directMis, (void *)rs->rayDiff, (void *)rs->dirRec,
                    (void *)rs->stackRec})
        if (p) cudaFree(p);
    rs->stackO = rs->stackD = rs->stackBeta = rs->shadowContrib = nullptr;
    rs->rayDiff = rs->dirRec = rs->stackRec = nullptr;
    rs->directShadow = nullptr;
    rs->directMis = nullptr;
    rs->directPaths = rs->directItems = 0;
}

// directlighting: kDirectStack pending vertices per path, `items` shadow / MIS rays per wavefront step
cudaError_t ensureDirectBuffers(RenderState *rs, size_t paths, size_t items, bool textured) {
    if (rs->directPaths >= paths && rs->directItems >= items && (!textured || rs->dirRec)) return cudaSuccess;
    paths = std::max(paths, rs->directPaths);
    items = std::max(items, rs->directItems);
    freeDirectBuffers(rs);
    cudaError_t e = cudaMalloc((void **)&rs->stackO, paths * kDirectStack * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void **)&rs->stackD, paths * kDirectStack * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void **)&rs->stackBeta, paths * kDirectStack * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void **)&rs->shadowContrib, items * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void **)&rs->directShadow, items * sizeof(ShadowItem));
    if (e == cudaSuccess) e = cudaMalloc((void **)&rs->directMis, items * sizeof(MisItem));
    if (textured) {  // ray differentials through the specular recursion: 2.1 KB per path
        if (e == cudaSuccess) e = cudaMalloc((void **)&rs->rayDiff, paths * 4 * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc((void **)&rs->dirRec, paths * kDirRec * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc((void **)&rs->stackRec, paths * kDirRec * kDirectStack * sizeof(float4));
    }
    if (e != cudaSuccess) {
        freeDirectBuffers(rs);
        return e;
    }
    rs->directPaths = paths;
    rs->directItems = items;
    return cudaSuccess;
}

// Path state + queues for `capacity` paths in flight (≈ 200 B per path).
cudaError_t ensureCapacity(RenderState *rs, size_t capacity, bool textured = false) {
    if (rs->capacity >= capacity && (!textured || rs->st.texRecord)) return cudaSuccess;
    capacity = std::max(capacity, rs->capacity);
    rs->st.texRecord = nullptr;
    freePathBuffers(rs);
    cudaError_t e = cudaSuccess;
#define A(ptr, n) if (e == cudaSuccess) e = devAlloc(rs, &(ptr), (n))
    A(rs->st.pixel, capacity);
    A(rs->st.sample, capacity);
    A(rs->st.rng, capacity);
    A(rs->st.rayO, capacity);
    A(rs->st.rayD, capacity);
    A(rs->st.beta, capacity);
    A(rs->st.L, capacity);
    A(rs->st.hit, capacity);
    A(rs->st.pendS, capacity);
    A(rs->st.pendM, capacity);
    if (textured) A(rs->st.texRecord, 8 * capacity);
    A(rs->q.active[0], capacity);
    A(rs->q.active[1], capacity);
    A(rs->q.shadow, capacity);
    A(rs->q.mis, capacity);
    A(rs->q.done, capacity);
    for (int c = 0; c < kNumClasses; ++c) A(rs->q.byClass[c], capacity);
    A(rs->q.counters, kNumCounters);
    A(rs->q.stats, 4);
#if IILE_RAY_SORT
    A(rs->sortKeys, capacity);
    A(rs->sortTmp, capacity);
    A(rs->sortHist, (size_t)kSortBins);
#endif
#undef A
    if (e != cudaSuccess) {
        freePathBuffers(rs);
        return e;
    }
    rs->capacity = capacity;
    return cudaSuccess;
}

// RadicalInverse(baseIndex, a) for the first five primes (lowdiscrepancy.cpp:388-403, 427-444)
float radicalInverse(int baseIndex, uint64_t a) {
    static const int primes[5] = {2, 3, 5, 7, 11};
    if (baseIndex == 0) {
        uint64_t r = 0, v = a;
        for (int i = 0; i < 64; ++i, v >>= 1) r = (r << 1) | (v & 1);
        return (float)((double)r * 0x1p-64);
    }
    const int base = primes[baseIndex];
    const float invBase = (float)1 / (float)base;
    uint64_t reversed = 0;
    float invBaseN = 1;
    while (a) {
        uint64_t next = a / base, digit = a - next * base;
        reversed = reversed * base + digit;
        invBaseN *= invBase;
        a = next;
    }
    return std::min((float)reversed * invBaseN, 0x1.fffffep-1f);
}

// HaltonSampler's static tables (halton.cpp:66-70): the first 1000 primes with their running
// sums (lowdiscrepancy.cpp Primes / PrimeSums) and ComputeRadicalInversePermutations(RNG())
// (lowdiscrepancy.cpp:2490-2504, Shuffle: sampling.h:150-157, RNG::UniformUInt32(b): rng.h:77-83).
cudaError_t ensureHaltonTables(RenderState *rs) {
    if (rs->haltonTables) return cudaSuccess;
    std::vector<uint32_t> primes;
    {
        const int limit = 8000;  // the 1000th prime is 7919
        std::vector<char> composite(limit, 0);
        for (int i = 2; i < limit && (int)primes.size() < kHaltonDims; ++i) {
            if (composite[i]) continue;
            primes.push_back((uint32_t)i);
            for (int j = i * i; j < limit; j += i) composite[j] = 1;
        }
    }
    HaltonTables host{};
    uint32_t sum = 0;
    for (int i = 0; i < kHaltonDims; ++i) {
        host.dimInfo[i] = make_uint2(primes[i], sum);
        sum += primes[i];
    }
    std::vector<uint16_t> perms(sum);
    Pcg32 rng;  // RNG(): PCG32_DEFAULT_STATE / PCG32_DEFAULT_STREAM (rng.h:60-61,94)
    rng.state = 0x853c49e6748fea9bULL;
    rng.inc = 0xda3e39cb94b95bdbULL;
    uint16_t *p = perms.data();
    for (int i = 0; i < kHaltonDims; ++i) {
        const int count = (int)primes[i];
        for (int j = 0; j < count; ++j) p[j] = (uint16_t)j;
        for (int j = 0; j < count; ++j) {
            const uint32_t b = (uint32_t)(count - j);
            const uint32_t threshold = (~b + 1u) % b;
            uint32_t r;
            do r = rng.uniformUInt32(); while (r < threshold);
            std::swap(p[j], p[j + (int)(r % b)]);
        }
        p += count;
    }
    cudaError_t e = cudaMalloc((void **)&rs->haltonPerms, perms.size() * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMemcpy(rs->haltonPerms, perms.data(), perms.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    host.perms = rs->haltonPerms;
    if (e == cudaSuccess) e = cudaMalloc((void **)&rs->haltonTables, sizeof(HaltonTables));
    if (e == cudaSuccess) e = cudaMemcpy(rs->haltonTables, &host, sizeof(host), cudaMemcpyHostToDevice);
    return e;
}

// SobolSampler ctor (sobol.h:64-80): resolution = RoundUpPow2(max extent of the sample bounds); the
// generator matrices and the two VdC rows SobolIntervalToIndex reads for that resolution go to the device.
int ensureSobol(RenderState *rs, const int sampleBounds[4], SobolConsts *out) {
    const SobolTables &t = sobolTables();
    int res = 1, log2Res = 0;
    const int extent = std::max(sampleBounds[2] - sampleBounds[0], sampleBounds[3] - sampleBounds[1]);
    while (res < extent) {
        res <<= 1;
        ++log2Res;
    }
    if (log2Res > t.nVdc)
        return setError(IILE_E_UNSUPPORTED, "Sampler \"sobol\": the film is larger than the 2^25 pixels a side "
                                            "SobolIntervalToIndex has matrices for");
    cudaError_t e = cudaSuccess;
    if (!rs->sobolMatrices) {
        const size_t bytes = (size_t)t.nDims * t.matrixSize * sizeof(uint32_t);
        e = cudaMalloc((void **)&rs->sobolMatrices, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(rs->sobolMatrices, t.matrices32, bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc((void **)&rs->sobolVdc, 2 * kSobolMatrixSize * sizeof(uint64_t));
    }
    if (e == cudaSuccess && rs->sobolVdcLog2 != log2Res) {
        uint64_t rows[2 * kSobolMatrixSize] = {};
        if (log2Res > 0) {
            std::memcpy(rows, t.vdc + (size_t)(log2Res - 1) * t.matrixSize, kSobolMatrixSize * sizeof(uint64_t));
            std::memcpy(rows + kSobolMatrixSize, t.vdcInv + (size_t)(log2Res - 1) * t.matrixSize,
                        kSobolMatrixSize * sizeof(uint64_t));
        }
        e = cudaMemcpy(rs->sobolVdc, rows, sizeof(rows), cudaMemcpyHostToDevice);
        rs->sobolVdcLog2 = log2Res;
    }
    if (e != cudaSuccess) return cudaFail(e, "sobol tables upload");
    out->matrices = rs->sobolMatrices;
    out->vdc = rs->sobolVdc;
    out->log2Resolution = log2Res;
    out->resolution = res;
    return IILE_OK;
}

// HaltonSampler ctor (halton.cpp:63-91): base-2 / base-3 scales covering min(res, 128) pixels and
// the multiplicative inverses used by GetIndexForSample.
void haltonConsts(const IileRenderSetup &S, const int sampleBounds[4], const HaltonTables *tables, HaltonConsts *h) {
    h->tables = tables;
    h->samplePixelCenter = S.samplePixelCenter;
    const int res[2] = {sampleBounds[2] - sampleBounds[0], sampleBounds[3] - sampleBounds[1]};
    for (int i = 0; i < 2; ++i) {
        const int base = i == 0 ? 2 : 3;
        int scale = 1, exp = 0;
        while (scale < std::min(res[i], kHaltonMaxResolution)) {
            scale *= base;
            ++exp;
        }
        h->baseScales[i] = scale;
        h->baseExponents[i] = exp;
    }
    h->sampleStride = h->baseScales[0] * h->baseScales[1];
    // multiplicativeInverse(a, n) = x mod n with a*x = 1 (mod n); n <= 243, found by search
    auto inverseMod = [](int64_t a, int64_t n) -> int {
        for (int64_t x = 0; x < n; ++x)
            if ((a * x) % n == 1 % n) return (int)x;
        return 0;
    };
    h->multInverse[0] = inverseMod(h->baseScales[1], h->baseScales[0]);
    h->multInverse[1] = inverseMod(h->baseScales[0], h->baseScales[1]);
}

// Distribution1D ctor (sampling.h:57-69) laid out as {cdf[n+1], func[n], funcInt}
void distribution1D(const std::vector<float> &f, std::vector<float> *out) {
    const int n = (int)f.size();
    out->assign(2 * (size_t)n + 2, 0.f);
    float *cdf = out->data(), *func = cdf + n + 1;
    for (int i = 0; i < n; ++i) func[i] = f[i];
    cdf[0] = 0;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / n;
    float funcInt = cdf[n];
    if (funcInt == 0)
        for (int i = 1; i < n + 1; ++i) cdf[i] = (float)i / (float)n;
    else
        for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    cdf[2 * n + 1] = funcInt;
}

int ensureState(IileScene *scene) {
    if (scene->renderState) return IILE_OK;
    RenderState *rs = new RenderState();
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&rs->smCount, cudaDevAttrMultiProcessorCount, dev);
    if (rs->smCount <= 0) rs->smCount = 148;
    cudaError_t e = cudaEventCreate(&rs->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&rs->ev1);
    if (e != cudaSuccess) {
        delete rs;
        return cudaFail(e, "cudaEventCreate");
    }
    scene->renderState = rs;
    return IILE_OK;
}

// CreateLightSampleDistribution (lightdistrib.cpp:47-66): one light or "uniform" -> uniform;
// "power" -> Light::Power().y() (diffuse.cpp:64-66, point.cpp:55); "spatial" -> the voxel grid,
// every voxel computed up front instead of lazily (the values do not depend on the order).
int ensureLightTables(IileScene *scene, int wantStrategy = -1) {
    RenderState *rs = stateOf(scene);
    if (wantStrategy < 0) wantStrategy = scene->host.setup.lightSampleStrategy;
    if (scene->host.lights.size() <= 1) wantStrategy = IILE_LIGHTSTRATEGY_UNIFORM;
    if (scene->lightTablesReady && rs->tableStrategy == wantStrategy) return IILE_OK;
    const HostScene &h = scene->host;
    DeviceScene &d = scene->dev;
    if (rs->lightTable) {
        cudaFree(rs->lightTable);
        rs->lightTable = nullptr;
    }
    d.lightCdf = nullptr;
    const int n = (int)h.lights.size();
    const int strategy = wantStrategy;
    rs->tableStrategy = strategy;
    d.lightStrategy = strategy;
    d.voxels[0] = d.voxels[1] = d.voxels[2] = 1;
    if (n == 0) {
        scene->lightTablesReady = true;
        return IILE_OK;
    }
    cudaError_t e;
    if (strategy != IILE_LIGHTSTRATEGY_SPATIAL) {
        std::vector<float> f(n, 1.f), table;
        if (strategy == IILE_LIGHTSTRATEGY_POWER) {
            const float pi = 3.14159265358979323846f;
            for (int i = 0; i < n; ++i) {
                const IileLight &L = h.lights[i];
                float p[3];
                for (int c = 0; c < 3; ++c) {
                    // Light::Power(): point.cpp:55, spot.cpp:77-79, distant.cpp:63-65, diffuse.cpp:59-61
                    if (L.type == IILE_LIGHT_POINT) p[c] = 4 * pi * L.L[c];
                    else if (L.type == IILE_LIGHT_SPOT)
                        p[c] = L.L[c] * 2 * pi * (1 - .5f * (L.cosFalloffStart + L.cosTotalWidth));
                    else if (L.type == IILE_LIGHT_DISTANT) p[c] = L.L[c] * pi * L.worldRadius * L.worldRadius;
                    else if (L.type == IILE_LIGHT_INFINITE)  // infinite.cpp:86-90: Lookup((.5,.5), .5) is the texel itself
                        p[c] = pi * L.worldRadius * L.worldRadius * L.L[c];
                    else p[c] = (IILE_LIGHT_TWOSIDED(L) ? 2 : 1) * L.L[c] * L.area * pi;
                }
                f[i] = 0.212671f * p[0] + 0.715160f * p[1] + 0.072169f * p[2];
            }
        }
        distribution1D(f, &table);
        e = cudaMalloc((void **)&rs->lightTable, table.size() * sizeof(float));
        if (e == cudaSuccess)
            e = cudaMemcpy(rs->lightTable, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return cudaFail(e, "light table upload");
    } else {
        // SpatialLightDistribution ctor, lightdistrib.cpp:95-110
        V3 diag = h.worldBound.diagonal();
        float bmax = diag[h.worldBound.maximumExtent()];
        size_t nVox = 1;
        for (int i = 0; i < 3; ++i) {
            d.voxels[i] = std::max(1, (int)std::round(diag[i] / bmax * 64));
            nVox *= (size_t)d.voxels[i];
        }
        const size_t floats = nVox * (2 * (size_t)n + 2);
        if (floats * sizeof(float) > ((size_t)48 << 30))
            return setError(IILE_E_UNSUPPORTED,
                            "spatial light distribution: dense voxel table would exceed 48 GiB; use "
                            "\"lightsamplestrategy\" \"power\" or \"uniform\" for this many lights");
        std::vector<float> halton(128 * 5);
        for (int i = 0; i < 128; ++i)
            for (int b = 0; b < 5; ++b) halton[5 * i + b] = radicalInverse(b, (uint64_t)i);
        float *dHalton = nullptr;
        e = cudaMalloc((void **)&rs->lightTable, floats * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void **)&dHalton, halton.size() * sizeof(float));
        if (e == cudaSuccess)
            e = cudaMemcpy(dHalton, halton.data(), halton.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = buildSpatialLightGrid(d, dHalton, rs->lightTable, 0);
        cudaFree(dHalton);
        if (e != cudaSuccess) return cudaFail(e, "spatial light grid build");
    }
    d.lightCdf = rs->lightTable;
    scene->lightTablesReady = true;
    return IILE_OK;
}

void copyM(const M44 &m, float *out) { std::memcpy(out, m.m, 16 * sizeof(float)); }

// ProjectiveCamera / PerspectiveCamera ctor (camera.h:95-110, perspective.cpp:50-57) and Film
void fillConsts(const IileRenderSetup &S, RenderConsts *rc) {
    std::memset(rc, 0, sizeof(*rc));
    Xform s2r = xfScale((float)S.xRes, (float)S.yRes, 1) *
                xfScale(1 / (S.screenWindow[1] - S.screenWindow[0]), 1 / (S.screenWindow[2] - S.screenWindow[3]), 1) *
                xfTranslate(V3(-S.screenWindow[0], -S.screenWindow[3], 0));
    Xform r2s = inverse(s2r);
    // PerspectiveCamera: Perspective(fov, 1e-2, 1000) (perspective.cpp:53); OrthographicCamera:
    // Orthographic(0, 1) = Scale(1, 1, 1 / (zFar - zNear)) * Translate(0, 0, -zNear) (orthographic.cpp:51, transform.cpp:293-295)
    Xform c2s = S.cameraType == IILE_CAMERA_ORTHOGRAPHIC ? xfScale(1, 1, 1 / (1.f - 0.f)) * xfTranslate(V3(0, 0, -0.f))
                                                         : xfPerspective(S.fov, 1e-2f, 1000.f);
    rc->cameraType = S.cameraType;
    rc->fullRes[0] = (float)S.xRes;
    rc->fullRes[1] = (float)S.yRes;
    Xform r2c = inverse(c2s) * r2s;
    copyM(r2c.m, rc->rasterToCamera);
    std::memcpy(rc->cameraToWorld, S.cameraToWorld, sizeof(rc->cameraToWorld));
    std::memcpy(rc->worldToCamera, S.worldToCamera, sizeof(rc->worldToCamera));
    rc->lensRadius = S.lensRadius;
    rc->focalDistance = S.focalDistance;
    buildFilterTable(S, rc->filterTable);
    for (int k = 0; k < 2; ++k) {
        rc->filterRadius[k] = S.filterRadius[k];
        rc->invFilterRadius[k] = 1.f / S.filterRadius[k];
    }
    rc->maxSampleLuminance = S.maxSampleLuminance;
    rc->filmMonitor = S.filterType == IILE_FILTER_MONITOR;
    FilmGeometry fg = filmGeometry(S);
    std::memcpy(rc->sampleBounds, fg.sampleBounds, sizeof(fg.sampleBounds));
    std::memcpy(rc->cropped, fg.cropped, sizeof(fg.cropped));
    std::memcpy(rc->pixelBounds, fg.sampleBounds, sizeof(fg.sampleBounds));
    if (S.hasPixelBounds) {  // path.cpp:216-228: x0 x1 y0 y1, intersected with the sample bounds
        rc->pixelBounds[0] = std::max(S.pixelBounds[0], fg.sampleBounds[0]);
        rc->pixelBounds[1] = std::max(S.pixelBounds[2], fg.sampleBounds[1]);
        rc->pixelBounds[2] = std::min(S.pixelBounds[1], fg.sampleBounds[2]);
        rc->pixelBounds[3] = std::min(S.pixelBounds[3], fg.sampleBounds[3]);
    }
    rc->spp = S.pixelSamples;
    rc->firstDimension = 5;
    rc->maxDepth = S.maxDepth;
    rc->rrThreshold = S.rrThreshold;
    rc->tilesX = (fg.sampleBounds[2] - fg.sampleBounds[0] + 15) / 16;  // integrator.cpp:237-240
    rc->tilesY = (fg.sampleBounds[3] - fg.sampleBounds[1] + 15) / 16;
}

struct Driver {
    IileScene *scene;
    RenderState *rs;
    cudaStream_t stream;
    DeviceScene dev{};  // the scene's device view with this stream's traversal counters (deviceViewFor)
    uint64_t launches = 0;
    cudaError_t err = cudaSuccess;
    // kernel-family timing (IILE_RENDER_TIME_KERNELS): an event is recorded in front of every launch;
    // the time between two consecutive events belongs to the family of the launch between them
    enum : int { kCast = 0, kShade = 1, kOther = 2 };
    bool timing = false;
    std::vector<int> marks;  // family of the launch that follows event i
    WavefrontLaunch wl() const { return WavefrontLaunch{stream, rs->smCount * 8}; }
    void mark(int family) {
        if (!timing || err != cudaSuccess) return;
        if (marks.size() >= rs->timingEvents.size()) {
            cudaEvent_t ev = nullptr;
            err = cudaEventCreate(&ev);
            if (err != cudaSuccess) return;
            rs->timingEvents.push_back(ev);
        }
        err = cudaEventRecord(rs->timingEvents[marks.size()], stream);
        marks.push_back(family);
    }
    void ck(cudaError_t e) {
        ++launches;
        if (err == cudaSuccess) err = e;
    }
    // synchronises; seconds[family] += the time of that family's launches
    cudaError_t readTiming(double seconds[3]) {
        seconds[0] = seconds[1] = seconds[2] = 0;
        if (!timing || marks.empty()) return cudaSuccess;
        mark(kOther);  // closing event
        cudaError_t e = err == cudaSuccess ? cudaStreamSynchronize(stream) : err;
        for (size_t i = 0; i + 1 < marks.size() && e == cudaSuccess; ++i) {
            float ms = 0;
            e = cudaEventElapsedTime(&ms, rs->timingEvents[i], rs->timingEvents[i + 1]);
            seconds[marks[i]] += ms * 1e-3;
        }
        return e;
    }
    // IILE_SHADE_FAST=0 sends every class through the generic kernel (A/B switch)
    void shade(const WavefrontLaunch &l, const RenderConsts &rc, int cur, HemiOutputs hemi) {
        static const bool fastOn = [] {
            const char *v = std::getenv("IILE_SHADE_FAST");
            return !v || std::atoi(v) != 0;
        }();
        // per-material kernels: scenes whose lights they sample in line (with a sphere light or an
        // infinite light the out-of-line sampler would be called from them, and the generic kernel is the
        // better one: cfg1 438 vs 420 Msamples/s, profiles/r7e_ab_shade_fast.txt)
        const int present = scene->classesPresent;
        const int fast = (fastOn && scene->simpleLights)
                             ? present & ((1 << kClassMiss) | (1 << kClassMatte) | (1 << kClassPlastic)) : 0;
        const cudaError_t e = wfShade(l, dev, rc, rs->st, rs->q, cur, hemi, present & ~fast, fast, scene->simpleLights, &launches);
        if (err == cudaSuccess) err = e;
    }
    void cast(const WavefrontLaunch &l, PathState st, Queues q, int cur) {
        if (rs->dump.rays) {
            mark(kOther);
            ++rs->dump.phase;
            ck(wfDumpRays(l, st, q, cur, rs->dump));
        }
        mark(kCast);
        ck(wfCast(l, dev, st, q, cur));
    }
    // paths still in the active queue `cur` (synchronises the stream)
    uint32_t activeCount(int cur) {
        uint32_t n = 0;
        if (err == cudaSuccess) err = cudaMemcpyAsync(&n, rs->q.counters + cur, sizeof(n), cudaMemcpyDeviceToHost, stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
        return err == cudaSuccess ? n : 0;
    }
    // One batch of `nWork` paths: generate, then maxDepth+1 bounces.  The rays a shade emits for
    // EstimateDirect (shadow, MIS) travel with the next bounce's closest-hit rays in one cast
    // launch; paths that ended are filmed after the cast that resolved their last direct light.
    void batch(const RenderConsts &rc, uint64_t firstWork, uint32_t nWork, HemiOutputs hemi, float4 *film) {
        if (err == cudaSuccess) err = cudaMemsetAsync(rs->q.counters, 0, kNumCounters * sizeof(uint32_t), stream);
        WavefrontLaunch l = wl();
        mark(kOther), ck(wfGenerate(l, rc, rs->st, rs->q, firstWork, nWork, hemi));
        int cur = 0;
        auto step = [&](int depth) {
#if IILE_RAY_SORT
            // camera rays leave the generator in tile order; from the first bounce on the queue is
            // binned by origin cell and direction octant before it is cast
            if (depth >= 1) {
                ck(wfSortActive(l, dev, rs->st, rs->q, cur, rs->sortKeys, rs->sortHist, rs->sortTmp));
                launches += 3;
            }
#endif
            cast(l, rs->st, rs->q, cur);
            if (depth > 0 && dev.nLights > 0) mark(kOther), ck(wfMisResolve(l, dev, rs->st, rs->q));
            if (depth > 0) mark(kOther), ck(wfFilm(l, rc, rs->st, rs->q, film, nullptr));
            mark(kOther), ck(wfResetCounters(l, rs->q, cur ^ 1));
            mark(kShade), ck(wfClassify(l, dev, rs->st, rs->q, cur));
            if (rs->st.texRecord && !scene->host.materialTextures.empty()) mark(kShade), ck(wfTextures(l, dev, rc, rs->st, rs->q, hemi));
            mark(kShade), shade(l, rc, cur, hemi);
            cur ^= 1;
        };
        for (int depth = 0; depth <= rc.maxDepth && err == cudaSuccess; ++depth) {
            step(depth);
            // a deep maxdepth with Russian roulette: the paths are gone long before the last step (as batchDirect)
            if (rc.maxDepth >= 12 && depth >= 8 && (depth & 3) == 3 && depth < rc.maxDepth && activeCount(cur) == 0) break;
        }
        // crossing an interface-only primitive takes a step but counts no bounce (path.cpp:111-116): such paths
        // still have vertices to visit — go on until the queue is empty (the host reads its length; scenes with
        // such primitives only)
        if (scene->hasInterfaces)
            for (int extra = 0; extra < kMaxInterfaceSteps && err == cudaSuccess && activeCount(cur) > 0; ++extra)
                step(rc.maxDepth + 1 + extra);
        // the last shade emits rays only for paths that crossed interface-only primitives
        // (bounces < maxDepth after maxDepth+1 vertices); resolve those, then film what is left
        if (dev.nLights > 0) {
            cast(l, rs->st, rs->q, -1);
            mark(kOther), ck(wfMisResolve(l, dev, rs->st, rs->q));
        }
        mark(kOther), ck(wfFilm(l, rc, rs->st, rs->q, film, nullptr));
    }
    // Integrator "directlighting": every wavefront step advances each path by one vertex of its
    // depth-first walk over the specular recursion; a camera ray has at most 2^maxDepth - 1 of them.
    void batchDirect(const RenderConsts &rc, uint64_t firstWork, uint32_t nWork, float4 *film, int strategy,
                     bool laned = false) {
        if (err == cudaSuccess) err = cudaMemsetAsync(rs->q.counters, 0, kNumCounters * sizeof(uint32_t), stream);
        WavefrontLaunch l = wl();
        PathState st = rs->st;
        st.stackO = rs->stackO;
        st.stackD = rs->stackD;
        st.stackBeta = rs->stackBeta;
        const bool textured = !scene->host.materialTextures.empty();
        st.rayDiff = textured ? rs->rayDiff : nullptr;
        st.dirRec = textured ? rs->dirRec : nullptr;
        st.stackRec = textured ? rs->stackRec : nullptr;
        Queues q = rs->q;
        q.shadow = rs->directShadow;
        q.mis = rs->directMis;
        q.shadowContrib = rs->shadowContrib;
        q.directMode = 1;
        mark(kOther), ck(wfGenerate(l, rc, st, q, firstWork, nWork, HemiOutputs{}));
        const int steps = (1 << std::min(std::max(rc.maxDepth, 1), kDirectStack + 1)) - 1;
        int cur = 0;
        auto step = [&](int it) {
            cast(l, st, q, cur);
            if (it > 0 && dev.nLights > 0) mark(kOther), ck(wfMisResolve(l, dev, st, q));
            if (it > 0) mark(kOther), ck(wfFilm(l, rc, st, q, film, nullptr));
            mark(kOther), ck(wfResetCounters(l, q, cur ^ 1));
            if (st.dirRec) mark(kShade), ck(wfTexturesDirect(l, dev, rc, st, q, cur));
            mark(kShade), ck(wfShadeDirect(l, dev, rc, st, q, cur, strategy));
            cur ^= 1;
        };
        for (int it = 0; it < steps && err == cudaSuccess; ++it) {
            step(it);
            // Most paths end at their first vertex (only specular surfaces recurse): once the queue is empty the
            // remaining steps — up to 2^maxdepth - 1 of them — would be empty launches.  One read-back after the second
            // step and then every fourth (cfg3d: 157 -> 12 launches per batch, 130 -> 148 Msamples/s; cfg1d 134 -> 149:
            // profiles/r9_direct_early_out.txt).
            if (!laned && steps > 4 && it + 1 < steps && (it == 1 || (it & 3) == 3) && activeCount(cur) == 0) break;
        }
        // Li(isect.SpawnRay(ray.d), ..., depth) behind an interface-only primitive (directlighting.cpp:74-75) is a
        // step that visits no vertex: as in batch()
        if (scene->hasInterfaces)
            for (int extra = 0; extra < kMaxInterfaceSteps && err == cudaSuccess && activeCount(cur) > 0; ++extra)
                step(steps + extra);
        if (dev.nLights > 0) {
            cast(l, st, q, -1);
            mark(kOther), ck(wfMisResolve(l, dev, st, q));
        }
        mark(kOther), ck(wfFilm(l, rc, st, q, film, nullptr));
    }
};

// IILE_RENDER_LANES: batches in flight at once (1 = strictly one after the other).  Default 4, and no
// more than one lane per 2^20 paths of the render (profiles/r7c_sweep_lanes.txt: cfg5 507 -> 587, cfg1 at
// 256 spp 461 -> 535, cfg3 201 -> 224 Msamples/s with four lanes; the 2.56 M paths of cfg1 at 16 spp are
// best cut in two, 413 -> 420, and lose when cut in four, 400).
int renderLanes(uint64_t totalWork) {
    static const int n = [] {
        const char *v = std::getenv("IILE_RENDER_LANES");
        const int k = v ? std::atoi(v) : kMaxLanes;
        return std::min(std::max(k, 1), kMaxLanes);
    }();
    return (int)std::min<uint64_t>((uint64_t)n, std::max<uint64_t>(totalWork >> 20, 1));
}

size_t pickCapacity(uint64_t totalWork, uint64_t requested) {
    static const uint64_t envCap = [] {  // IILE_PATHS_IN_FLIGHT: tuning override of the default batch size
        const char *v = std::getenv("IILE_PATHS_IN_FLIGHT");
        return v ? (uint64_t)std::strtoull(v, nullptr, 10) : (uint64_t)0;
    }();
    // default 2^24 paths per batch and lane (3.4 GB of path state and queues each): against 2^22, cfg5 at 32 spp
    // 687 -> 716 and cfg3 at 32 spp 252 -> 259 Msamples/s, a quarter of the launches (profiles/r8g_sweep_batch.txt)
    uint64_t cap = requested ? requested : (envCap ? envCap : ((uint64_t)1 << 24));
    cap = std::min<uint64_t>(cap, (uint64_t)1 << 26);
    cap = std::max<uint64_t>(std::min(cap, totalWork), 1024);
    return (size_t)((cap + 255) & ~(uint64_t)255);
}

int readStats(Driver &dr, IileRenderStats *stats, bool sync) {
    if (!stats) return IILE_OK;
    std::memset(stats, 0, sizeof(*stats));
    stats->launches = dr.launches;
    if (!sync) return IILE_OK;
    unsigned long long h[kMaxLanes][4] = {};
    cudaError_t e = cudaSuccess;
    const int lanes = dr.rs->lanesUsed;
    for (int k = 0; k < lanes && e == cudaSuccess; ++k)  // the lanes have joined dr.stream by now
        e = cudaMemcpyAsync(h[k], (k ? dr.rs->aux[k - 1] : dr.rs)->q.stats, sizeof(h[k]), cudaMemcpyDeviceToHost, dr.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dr.stream);
    if (e != cudaSuccess) return cudaFail(e, "render statistics read-back");
    for (int k = 0; k < lanes; ++k) {
        stats->closestRays += h[k][0];
        stats->shadowRays += h[k][1];
        stats->cameraRays += h[k][2];
    }
    return IILE_OK;
}

int prepare(IileScene *scene, const char *who, int lightStrategy = -1) {
    if (!scene) return setError(IILE_E_INVALID, std::string(who) + ": null scene");
    if (!scene->uploaded) {
        int rc = uploadScene(scene);
        if (rc != IILE_OK) return rc;
    }
    int rc = ensureState(scene);
    if (rc != IILE_OK) return rc;
    return ensureLightTables(scene, lightStrategy);
}

// CreateHemisphericCamera's LookAt and its numerically re-inverted WorldToCamera (hemispheric.cpp:117-134)
void hemiCameraMatrices(const float *pos, const float *dir, M44 *c2w, M44 *w2c, M44 *w2cInv) {
    const V3 p(pos[0], pos[1], pos[2]), d(dir[0], dir[1], dir[2]);
    const V3 up = (d.x == 0.f && d.y == 0.f) ? V3(0, 1, 0) : V3(0, 0, 1);
    const Xform look = xfLookAt(p, V3(p.x + d.x, p.y + d.y, p.z + d.z), up);
    *c2w = look.inv;          // Transform(LookAt(...).GetInverseMatrix()).m
    *w2c = inverse(*c2w);     // WorldToCamera.m
    *w2cInv = inverse(*w2c);  // WorldToCamera.mInv, what Normal3f uses
}

// One batch of hemispheres enqueued on a stream; the maps stay on the device until release().
struct HemiBatch {
    float4 *film = nullptr;
    float *normals = nullptr, *distance = nullptr, *cams = nullptr;
    size_t nPix = 0;
    uint64_t launches = 0;
    IileRenderSetup setup{};
    void release(cudaStream_t stream) {
        for (void *p : {(void *)film, (void *)normals, (void *)distance, (void *)cams})
            if (p) cudaFreeAsync(p, stream);
        film = nullptr, normals = distance = cams = nullptr;
    }
};

int hemiEnqueue(IileScene *scene, const char *who, size_t n, const float *pos, const float *dir, int size, int spp,
                cudaStream_t stream, bool timed, HemiBatch *hb) {
    int rc = prepare(scene, who, IILE_LIGHTSTRATEGY_SPATIAL);
    if (rc != IILE_OK) return rc;
    if (!pos || !dir) return setError(IILE_E_INVALID, std::string(who) + ": null buffer");
    if (size <= 0 || size > 1024 || spp <= 0)
        return setError(IILE_E_INVALID, std::string(who) + ": size / spp out of range");
    if (n * (size_t)spp >= ((size_t)1 << 32))
        return setError(IILE_E_INVALID, std::string(who) + ": too many (hemisphere, sample) streams in one batch");
    RenderState *rs = stateOf(scene);
    // per-hemisphere camera: CreateHemisphericCamera's LookAt and its numerically re-inverted
    // WorldToCamera (hemispheric.cpp:120-137)
    std::vector<float> cams(32 * n);
    try {
        for (size_t h = 0; h < n; ++h) {
            M44 c2w, w2c, w2cInv;
            hemiCameraMatrices(pos + 3 * h, dir + 3 * h, &c2w, &w2c, &w2cInv);
            copyM(c2w, &cams[32 * h]);
            copyM(w2cInv, &cams[32 * h + 16]);
        }
    } catch (const std::exception &ex) {
        return setError(IILE_E_INVALID, std::string(who) + ": " + ex.what());
    }
    IileRenderSetup S = scene->host.setup;
    S.xRes = S.yRes = size;
    S.cropWindow[0] = S.cropWindow[2] = 0.f;
    S.cropWindow[1] = S.cropWindow[3] = 1.f;
    S.filterType = IILE_FILTER_GAUSSIAN;
    S.filterRadius[0] = S.filterRadius[1] = 2.f;
    S.filterParam[0] = 2.f;
    S.filmScale = 1.f;
    S.maxSampleLuminance = std::numeric_limits<float>::infinity();
    S.pixelSamples = spp;
    S.maxDepth = 3;        // CreateIISPTdIntegrator, iispt_d.cpp:505
    S.rrThreshold = 0.5f;  // iispt_d.h:73
    S.hasPixelBounds = 0;
    hb->setup = S;
    RenderConsts rcs;
    fillConsts(S, &rcs);
    rcs.hemiMode = 1;
    rcs.hemiSize = size;
    rcs.sampleBegin = 0;
    rcs.sampleEnd = spp;
    // only the pixels of [0,size)^2 are traced (pixelBounds); streams are indexed inside them
    rcs.sampleBounds[0] = rcs.sampleBounds[1] = rcs.pixelBounds[0] = rcs.pixelBounds[1] = 0;
    rcs.sampleBounds[2] = rcs.sampleBounds[3] = rcs.pixelBounds[2] = rcs.pixelBounds[3] = size;
    const size_t nPix = n * (size_t)size * size;
    hb->nPix = nPix;
    cudaError_t e = cudaMallocAsync((void **)&hb->film, nPix * sizeof(float4), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&hb->normals, nPix * 3 * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&hb->distance, nPix * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&hb->cams, cams.size() * sizeof(float), stream);
    // pageable source: the copy is staged before the call returns, `cams` may go out of scope
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(hb->cams, cams.data(), cams.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(hb->film, 0, nPix * sizeof(float4), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(hb->normals, 0, nPix * 3 * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(hb->distance, 0, nPix * sizeof(float), stream);
    Driver dr{scene, rs, stream};
    const uint64_t total = (uint64_t)nPix * (uint64_t)spp;
    const size_t cap = pickCapacity(total, 0);
    if ((rc = deviceViewFor(scene, stream, &dr.dev)) != IILE_OK) return rc;
    if (e == cudaSuccess) e = ensureCapacity(rs, cap, !scene->host.materialTextures.empty());
    rs->lanesUsed = 1;
    if (e == cudaSuccess) e = cudaMemsetAsync(rs->q.stats, 0, 4 * sizeof(unsigned long long), dr.stream);
    if (e == cudaSuccess && timed) e = cudaEventRecord(rs->ev0, dr.stream);
    if (e == cudaSuccess) {
        HemiOutputs hemi{hb->cams, hb->normals, hb->distance};
        for (uint64_t first = 0; first < total && dr.err == cudaSuccess; first += cap)
            dr.batch(rcs, first, (uint32_t)std::min<uint64_t>(cap, total - first), hemi, hb->film);
        e = dr.err;
    }
    if (e == cudaSuccess && timed) e = cudaEventRecord(rs->ev1, dr.stream);
    hb->launches = dr.launches;
    if (e != cudaSuccess) return cudaFail(e, who);
    return IILE_OK;
}

}  // namespace

IileScene::~IileScene() {
    freeRenderState(this);
    IileSceneFree(this);
}

namespace iile {
void freeRenderState(IileScene *s) {
    RenderState *rs = static_cast<RenderState *>(s->renderState);
    if (!rs) return;
    freePathBuffers(rs);
    if (rs->lightTable) cudaFree(rs->lightTable);
    if (rs->hemiCams) cudaFree(rs->hemiCams);
    if (rs->haltonTables) cudaFree(rs->haltonTables);
    if (rs->haltonPerms) cudaFree(rs->haltonPerms);
    if (rs->sobolMatrices) cudaFree(rs->sobolMatrices);
    if (rs->sobolVdc) cudaFree(rs->sobolVdc);
    if (rs->hostFilm) cudaFree(rs->hostFilm);
    freeDirectBuffers(rs);
    for (int b = 0; b < 2; ++b) {
        if (rs->pinned[b]) cudaFreeHost(rs->pinned[b]);
        if (rs->pinnedDone[b]) cudaEventDestroy(rs->pinnedDone[b]);
    }
    if (rs->copyStream) cudaStreamDestroy(rs->copyStream);
    for (cudaEvent_t ev : rs->timingEvents) cudaEventDestroy(ev);
    for (RenderState *a : rs->aux)
        if (a) {
            freePathBuffers(a);
            freeDirectBuffers(a);
            delete a;
        }
    for (int k = 0; k < kMaxLanes; ++k) {
        if (rs->laneStream[k]) cudaStreamDestroy(rs->laneStream[k]);
        if (rs->laneDone[k]) cudaEventDestroy(rs->laneDone[k]);
    }
    if (rs->forkEvent) cudaEventDestroy(rs->forkEvent);
    if (rs->ev0) cudaEventDestroy(rs->ev0);
    if (rs->ev1) cudaEventDestroy(rs->ev1);
    delete rs;
    s->renderState = nullptr;
    s->lightTablesReady = false;
}
}  // namespace iile

extern "C" {

// Enqueues the whole render on `stream`.  `stats` (optional) forces one stream synchronisation
// at the end to read the ray counters back; pass NULL to stay asynchronous.
int iile_render_dev(IileScene *scene, const IileRenderParams *params, float *filmRGBW,
                    IileRenderStats *stats, void *stream) {
    int rc = prepare(scene, "iile_render_dev");
    if (rc != IILE_OK) return rc;
    if (!filmRGBW) return setError(IILE_E_INVALID, "iile_render_dev: null film");
    IileRenderParams p{};
    if (params) p = *params;
    const IileRenderSetup &S = scene->host.setup;
    RenderConsts rcs;
    fillConsts(S, &rcs);
    rcs.sampleBegin = std::max(0, p.sampleBegin);
    rcs.sampleEnd = p.sampleEnd <= 0 ? S.pixelSamples : std::min(p.sampleEnd, S.pixelSamples);
    const int nTiles = rcs.tilesX * rcs.tilesY;
    rcs.tileBegin = std::max(0, p.tileBegin);
    rcs.tileEnd = p.tileEnd <= 0 ? nTiles : std::min(p.tileEnd, nTiles);
    if (rcs.sampleBegin > rcs.sampleEnd || rcs.tileBegin > rcs.tileEnd)
        return setError(IILE_E_INVALID, "iile_render_dev: empty or reversed sample / tile range");
    RenderState *rs = stateOf(scene);
    if (S.samplerType == IILE_SAMPLER_HALTON) {
        // GlobalSampler dimensions: 5 for the camera sample, then at most 8 per bounce
        if (5 + 8 * (int64_t)std::max(S.maxDepth, 0) > kHaltonDims)
            return setError(IILE_E_UNSUPPORTED, "Sampler \"halton\": maxdepth this large can run past the 1000 "
                            "dimensions HaltonSampler can sample (the reference aborts there, halton.h:71-76)");
        cudaError_t e = ensureHaltonTables(rs);
        if (e != cudaSuccess) return cudaFail(e, "halton tables upload");
        haltonConsts(S, rcs.sampleBounds, rs->haltonTables, &rcs.halton);
    } else if (S.samplerType == IILE_SAMPLER_SOBOL) {
        // 5 dimensions for the camera sample, then at most 8 per bounce; past 1023 the reference draws the
        // dimension itself at random (IILE's change, sobol.cpp:48-50), which no second run reproduces
        if (5 + 8 * (int64_t)std::max(S.maxDepth, 0) > kSobolDims)
            return setError(IILE_E_UNSUPPORTED, "Sampler \"sobol\": maxdepth this large can run past the 1024 "
                                                "Sobol' dimensions");
        if (S.pixelSamples & (S.pixelSamples - 1))
            return setError(IILE_E_INVALID, "Sampler \"sobol\": pixelsamples must be a power of two (the reference "
                                            "rounds up, sobol.h:64-70)");
        if ((rc = ensureSobol(rs, rcs.sampleBounds, &rcs.halton.sobol)) != IILE_OK) return rc;
    } else if (S.samplerType != IILE_SAMPLER_RANDOM)
        return setError(IILE_E_INVALID, "iile_render_dev: unknown samplerType");
    Driver dr{scene, rs, (cudaStream_t)stream};
    dr.timing = stats != nullptr && (p.flags & IILE_RENDER_TIME_KERNELS) != 0;
    const uint64_t sppLocal = (uint64_t)(rcs.sampleEnd - rcs.sampleBegin);
    const uint64_t total = (uint64_t)(rcs.tileEnd - rcs.tileBegin) * 256u * sppLocal;
    const size_t fw = (size_t)(rcs.cropped[2] - rcs.cropped[0]), fh = (size_t)(rcs.cropped[3] - rcs.cropped[1]);
    if (p.clearFilm) dr.err = cudaMemsetAsync(filmRGBW, 0, fw * fh * sizeof(float4), dr.stream);
    const bool direct = S.integratorType == IILE_INTEGRATOR_DIRECT || S.integratorType == IILE_INTEGRATOR_WHITTED;
    // shadeDirectKernel's light loop: 0 = UniformSampleAllLights, 1 = UniformSampleOneLight, 2 = Whitted's
    const int directStrategy = S.integratorType == IILE_INTEGRATOR_WHITTED ? 2 : S.directStrategy;
    if (direct) {
        if (S.maxDepth > kDirectStack + 1)
            return setError(IILE_E_UNSUPPORTED, "Integrator \"directlighting\" / \"whitted\": maxdepth above 9 is not "
                                                "supported (one pending vertex per level of the specular recursion)");
        // DirectLightingIntegrator::Preprocess (directlighting.cpp:44-59): two 2D arrays per light and level
        if (S.integratorType == IILE_INTEGRATOR_DIRECT && directStrategy == IILE_DIRECT_ALL)
            rcs.firstDimension = 5 + 4 * (int)scene->host.lights.size() * std::max(S.maxDepth, 0);
        if (S.samplerType != IILE_SAMPLER_RANDOM && rcs.firstDimension + 4 * std::max(S.maxDepth, 0) > kHaltonDims)
            return setError(IILE_E_UNSUPPORTED, "Integrator \"directlighting\": the sample arrays of this many lights and levels "
                                                "run past the dimensions the halton / sobol sampler can sample");
    } else if (S.integratorType != IILE_INTEGRATOR_PATH)
        return setError(IILE_E_INVALID, "iile_render_dev: unknown integratorType");
    if (total > 0) {
        size_t cap = pickCapacity(total, p.maxPathsInFlight);
        size_t perPath = 1;
        if (direct) {  // several lights per vertex: bound the ray queues, not only the paths
            perPath = 1;
            if (directStrategy != IILE_DIRECT_ONE) {  // "all": "samples" terms per light and vertex (sample arrays)
                perPath = 0;
                for (const IileLight &l : scene->host.lights) perPath += directStrategy == IILE_DIRECT_ALL ? IILE_LIGHT_NSAMPLES(l) : 1;
                perPath = std::max<size_t>(perPath, 1);
            }
            size_t maxPaths = std::max<size_t>(((size_t)1 << 24) / perPath, (size_t)1 << 16);
            if (!scene->host.materialTextures.empty()) maxPaths = std::min<size_t>(maxPaths, (size_t)1 << 20);  // 2.1 KB of differentials per path
            cap = std::min(cap, (maxPaths + 255) & ~(size_t)255);
        }
        // lanes: measurement modes and the direct integrators run their batches one after the other
        static const bool directLanes = [] {
            const char *v = std::getenv("IILE_DIRECT_LANES");
            return v && std::atoi(v) != 0;
        }();
        int nLanes = ((direct && !directLanes) || dr.timing || rs->dump.rays) ? 1 : renderLanes(total);
        if (nLanes > 1 && !p.maxPathsInFlight && total > ((uint64_t)1 << 20) && total < (uint64_t)cap * nLanes)
            cap = (size_t)((((total + nLanes - 1) / nLanes) + 255) & ~(uint64_t)255);  // a small render still fills every lane
        nLanes = (int)std::min<uint64_t>(nLanes, (total + cap - 1) / cap);
        if ((rc = deviceViewFor(scene, dr.stream, &dr.dev)) != IILE_OK) return rc;
        const bool textured = !scene->host.materialTextures.empty();
        cudaError_t e = ensureCapacity(rs, cap, textured && !direct);
        if (e == cudaSuccess && direct) e = ensureDirectBuffers(rs, cap, cap * perPath, textured);
        for (int k = 1; k < nLanes && e == cudaSuccess; ++k) {
            if (!rs->aux[k - 1]) {
                rs->aux[k - 1] = new RenderState();
                rs->aux[k - 1]->smCount = rs->smCount;
            }
            e = ensureCapacity(rs->aux[k - 1], cap, textured && !direct);
            if (e == cudaSuccess && direct) e = ensureDirectBuffers(rs->aux[k - 1], cap, cap * perPath, textured);
        }
        for (int k = 0; k < nLanes && nLanes > 1 && e == cudaSuccess; ++k) {
            if (!rs->laneStream[k]) e = cudaStreamCreateWithFlags(&rs->laneStream[k], cudaStreamNonBlocking);
            if (e == cudaSuccess && !rs->laneDone[k]) e = cudaEventCreateWithFlags(&rs->laneDone[k], cudaEventDisableTiming);
        }
        if (e == cudaSuccess && nLanes > 1 && !rs->forkEvent) e = cudaEventCreateWithFlags(&rs->forkEvent, cudaEventDisableTiming);
        if (e != cudaSuccess) return cudaFail(e, "path-state allocation");
        HemiOutputs hemi{};
        if (nLanes == 1) {
            if (dr.err == cudaSuccess) dr.err = cudaMemsetAsync(rs->q.stats, 0, 4 * sizeof(unsigned long long), dr.stream);
            for (uint64_t first = 0; first < total && dr.err == cudaSuccess; first += cap) {
                const uint32_t nWork = (uint32_t)std::min<uint64_t>(cap, total - first);
                if (direct) dr.batchDirect(rcs, first, nWork, (float4 *)filmRGBW, directStrategy);
                else dr.batch(rcs, first, nWork, hemi, (float4 *)filmRGBW);
            }
        } else {
            // fork: every lane starts after what the caller's stream holds so far (film clear included)
            if (dr.err == cudaSuccess) dr.err = cudaEventRecord(rs->forkEvent, dr.stream);
            Driver lane[kMaxLanes];
            for (int k = 0; k < nLanes && dr.err == cudaSuccess; ++k) {
                lane[k] = Driver{scene, k ? rs->aux[k - 1] : rs, rs->laneStream[k]};
                if ((rc = deviceViewFor(scene, lane[k].stream, &lane[k].dev)) != IILE_OK) return rc;
                dr.err = cudaStreamWaitEvent(lane[k].stream, rs->forkEvent, 0);
                if (dr.err == cudaSuccess)
                    dr.err = cudaMemsetAsync(lane[k].rs->q.stats, 0, 4 * sizeof(unsigned long long), lane[k].stream);
            }
            int k = 0;
            for (uint64_t first = 0; first < total && dr.err == cudaSuccess; first += cap, k = (k + 1) % nLanes) {
                const uint32_t nWork = (uint32_t)std::min<uint64_t>(cap, total - first);
                if (direct) lane[k].batchDirect(rcs, first, nWork, (float4 *)filmRGBW, directStrategy, true);
                else lane[k].batch(rcs, first, nWork, hemi, (float4 *)filmRGBW);
                if (lane[k].err != cudaSuccess) dr.err = lane[k].err;
            }
            // join: the caller's stream continues when every lane has drained
            for (k = 0; k < nLanes && dr.err == cudaSuccess; ++k) {
                dr.err = cudaEventRecord(rs->laneDone[k], lane[k].stream);
                if (dr.err == cudaSuccess) dr.err = cudaStreamWaitEvent(dr.stream, rs->laneDone[k], 0);
                dr.launches += lane[k].launches;
            }
        }
        rs->lanesUsed = nLanes;
    }
    if (dr.err != cudaSuccess) return cudaFail(dr.err, "iile_render_dev");
    rs->lastLaunches = dr.launches;
    rc = readStats(dr, stats, stats != nullptr);
    if (rc == IILE_OK && dr.timing) {
        double sec[3];
        cudaError_t e = dr.readTiming(sec);
        if (e != cudaSuccess) return cudaFail(e, "kernel timing read-back");
        stats->castSeconds = sec[0];
        stats->shadeSeconds = sec[1];
        stats->otherSeconds = sec[2];
    }
    return rc;
}

int iile_render_dump_rays_dev(IileScene *scene, const IileRenderParams *params, float *filmRGBW, IileRay *rays,
                              uint8_t *kinds, uint64_t cap, uint32_t stride, uint64_t *count, void *stream) {
    int rc = prepare(scene, "iile_render_dump_rays_dev");
    if (rc != IILE_OK) return rc;
    if (!rays || !kinds || !count || cap == 0)
        return setError(IILE_E_INVALID, "iile_render_dump_rays_dev: null buffer or zero capacity");
    RenderState *rs = stateOf(scene);
    unsigned long long *dCount = nullptr;
    cudaError_t e = cudaMalloc((void **)&dCount, sizeof(*dCount));
    if (e == cudaSuccess) e = cudaMemsetAsync(dCount, 0, sizeof(*dCount), (cudaStream_t)stream);
    if (e != cudaSuccess) {
        if (dCount) cudaFree(dCount);
        return cudaFail(e, "iile_render_dump_rays_dev");
    }
    rs->dump = RayDump{rays, kinds, cap, dCount, stride ? stride : 1u, 0u};
    rc = iile_render_dev(scene, params, filmRGBW, nullptr, stream);
    rs->dump = RayDump{};
    unsigned long long h = 0;
    if (rc == IILE_OK) {
        e = cudaMemcpyAsync(&h, dCount, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) rc = cudaFail(e, "iile_render_dump_rays_dev");
    }
    cudaFree(dCount);
    *count = h;
    return rc;
}

int iile_render(IileScene *scene, const IileRenderParams *params, float *filmRGBW, IileRenderStats *stats) {
    int rc = prepare(scene, "iile_render");
    if (rc != IILE_OK) return rc;
    if (!filmRGBW) return setError(IILE_E_INVALID, "iile_render: null film");
    FilmGeometry fg = filmGeometry(scene->host.setup);
    const size_t bytes = (size_t)fg.width() * fg.height() * sizeof(float4);
    RenderState *rs = stateOf(scene);
    cudaError_t e = cudaSuccess;
    if (rs->hostFilmBytes < bytes || !rs->hostFilm) {  // the device film lives as long as the scene
        if (rs->hostFilm) cudaFree(rs->hostFilm);
        rs->hostFilm = nullptr;
        rs->hostFilmBytes = 0;
        e = cudaMalloc((void **)&rs->hostFilm, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return cudaFail(e, "cudaMalloc(film)");
        rs->hostFilmBytes = std::max<size_t>(bytes, 16);
    }
    float *dFilm = rs->hostFilm;
    IileRenderParams p{};
    if (params) p = *params;
    if (!p.clearFilm) e = cudaMemcpy(dFilm, filmRGBW, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaEventRecord(rs->ev0, 0);
    IileRenderStats local{};
    if (e == cudaSuccess) {
        rc = iile_render_dev(scene, &p, dFilm, nullptr, nullptr);
        if (rc != IILE_OK) return rc;
        e = cudaEventRecord(rs->ev1, 0);
    }
    if (e == cudaSuccess) e = downloadPipelined(rs, filmRGBW, dFilm, bytes, rs->ev1);
    if (e != cudaSuccess) return cudaFail(e, "iile_render");
    if (stats) {
        Driver dr{scene, rs, (cudaStream_t)0};
        rc = readStats(dr, &local, true);
        if (rc != IILE_OK) return rc;
        float ms = 0;
        cudaEventElapsedTime(&ms, rs->ev0, rs->ev1);
        local.deviceSeconds = ms * 1e-3;
        local.launches = rs->lastLaunches;
        *stats = local;
    }
    return IILE_OK;
}

int iile_film_resolve(const IileScene *scene, const float *filmRGBW, float *rgb) {
    if (!scene || !filmRGBW || !rgb) return setError(IILE_E_INVALID, "iile_film_resolve: null argument");
    FilmGeometry g = filmGeometry(scene->host.setup);
    resolveFilm(scene->host.setup, filmRGBW, (size_t)g.width() * g.height(), rgb);
    return IILE_OK;
}

int iile_write_image(const char *path, const float *rgb, int xRes, int yRes, int fullX, int fullY,
                     int offX, int offY) {
    if (!path || !rgb) return setError(IILE_E_INVALID, "iile_write_image: null argument");
    try {
        writeImage(path, rgb, xRes, yRes, fullX, fullY, offX, offY);
    } catch (const std::exception &ex) {
        return setError(IILE_E_IO, ex.what());
    }
    return IILE_OK;
}

// IISPTdIntegrator::RenderView for a batch of hemispheric cameras (iispt_d.cpp:345-461,
// cameras/hemispheric.cpp:15-41,108-158): size x size film, Gaussian filter r = 2 alpha = 2,
// maxdepth 3, rrthreshold 0.5, spatial light distribution, no emission at bounce 0; the first
// hit's distance (-1 on a miss) and camera-space normal per pixel.  All hemispheres of the
// batch share the launches: work item = (hemisphere, pixel, sample).
int iile_render_hemispheres(IileScene *scene, size_t n, const float *pos, const float *dir, int size, int spp,
                            float *intensity, float *normals, float *distance, IileRenderStats *stats) {
    if (n == 0) return prepare(scene, "iile_render_hemispheres", IILE_LIGHTSTRATEGY_SPATIAL);
    if (!intensity || !normals || !distance)
        return setError(IILE_E_INVALID, "iile_render_hemispheres: null buffer");
    HemiBatch hb;
    int rc = hemiEnqueue(scene, "iile_render_hemispheres", n, pos, dir, size, spp, (cudaStream_t)0, true, &hb);
    cudaError_t e = cudaSuccess;
    if (rc == IILE_OK) {
        // Film::to_intensity_film -> to_rgb_array(1) on the device (the film's own storage is reused:
        // pixel i's RGB lands in floats [3i, 3i+3) of a separate buffer), then pipelined read-backs
        RenderState *rs = stateOf(scene);
        float *rgb = nullptr;
        e = cudaMallocAsync((void **)&rgb, hb.nPix * 3 * sizeof(float), (cudaStream_t)0);
        if (e == cudaSuccess) e = hemiResolve(hb.film, hb.nPix, hb.setup.filmScale, rgb, (cudaStream_t)0);
        if (e == cudaSuccess) e = cudaEventRecord(rs->ev1, (cudaStream_t)0);
        if (e == cudaSuccess) e = downloadPipelined(rs, intensity, rgb, hb.nPix * 3 * sizeof(float), rs->ev1);
        if (e == cudaSuccess) e = downloadPipelined(rs, normals, hb.normals, hb.nPix * 3 * sizeof(float), rs->ev1);
        if (e == cudaSuccess) e = downloadPipelined(rs, distance, hb.distance, hb.nPix * sizeof(float), rs->ev1);
        if (rgb) cudaFreeAsync(rgb, (cudaStream_t)0);
    }
    hb.release((cudaStream_t)0);
    if (rc != IILE_OK) return rc;
    if (e != cudaSuccess) return cudaFail(e, "iile_render_hemispheres");
    if (stats) {
        Driver dr{scene, stateOf(scene), (cudaStream_t)0};
        dr.launches = hb.launches;
        rc = readStats(dr, stats, true);
        if (rc != IILE_OK) return rc;
        float ms = 0;
        cudaEventElapsedTime(&ms, stateOf(scene)->ev0, stateOf(scene)->ev1);
        stats->deviceSeconds = ms * 1e-3;
    }
    return IILE_OK;
}

// RenderView + IisptRenderRunner::normalizeMapsDownstream (iisptrenderrunner.cpp:1041-1095) with
// everything left in HBM: nnInput [n][7][size][size] is the tensor main_stdio_net.read_input
// builds from what IisptNnConnector::communicate pipes (iisptnnconnector.cpp:125-148), means
// [n][3] the channel means transformMapsUpstream needs afterwards.  Asynchronous on `stream`
// unless `stats` is given (one synchronisation to read the ray counters).
int iile_hemispheres_to_nn_dev(IileScene *scene, size_t n, const float *pos, const float *dir, int size, int spp,
                               float *nnInput, float *means, IileRenderStats *stats, void *stream) {
    if (n == 0) return prepare(scene, "iile_hemispheres_to_nn_dev", IILE_LIGHTSTRATEGY_SPATIAL);
    if (!nnInput || !means) return setError(IILE_E_INVALID, "iile_hemispheres_to_nn_dev: null buffer");
    HemiBatch hb;
    int rc = hemiEnqueue(scene, "iile_hemispheres_to_nn_dev", n, pos, dir, size, spp, (cudaStream_t)stream, false, &hb);
    cudaError_t e = cudaSuccess;
    if (rc == IILE_OK)
        e = hemiToNN(hb.film, hb.normals, hb.distance, n, size, hb.setup.filmScale, nnInput, means, (cudaStream_t)stream);
    hb.release((cudaStream_t)stream);
    if (rc != IILE_OK) return rc;
    if (e != cudaSuccess) return cudaFail(e, "iile_hemispheres_to_nn_dev");
    if (stats) {
        Driver dr{scene, stateOf(scene), (cudaStream_t)stream};
        dr.launches = hb.launches + 1;
        return readStats(dr, stats, true);
    }
    return IILE_OK;
}

// IisptRenderRunner::transformMapsUpstream (iisptrenderrunner.cpp:1098-1136): the network's answer
// nnOutput [n][3][size][size] and the means of the matching iile_hemispheres_to_nn_dev call ->
// intensity [n][size][size][3] in ImageFilm order (IntensityFilm::populate_from_float_array).
int iile_hemispheres_from_nn_dev(size_t n, int size, const float *nnOutput, const float *means, float *intensity,
                                 void *stream) {
    if (n == 0) return IILE_OK;
    if (!nnOutput || !means || !intensity)
        return setError(IILE_E_INVALID, "iile_hemispheres_from_nn_dev: null buffer");
    if (size <= 0 || size > 1024) return setError(IILE_E_INVALID, "iile_hemispheres_from_nn_dev: size out of range");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev < 1)
        return setError(IILE_E_CUDA, "iile_hemispheres_from_nn_dev: no CUDA device (there is no CPU path)");
    cudaError_t e = hemiFromNN(nnOutput, means, n, size, intensity, (cudaStream_t)stream);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, "iile_hemispheres_from_nn_dev");
}
}

// ---- one task of IisptRenderRunner::run: the stages either side of the hemisphere batch (iispt_task.cu) ----
namespace {
int checkTask(const IileIisptTask *t, const char *who, IisptTaskDev *out) {
    if (!t) return setError(IILE_E_INVALID, std::string(who) + ": null task");
    if (t->x1 <= t->x0 || t->y1 <= t->y0 || t->tilesize <= 0 || t->hemiSize <= 0 || t->hemiSize > 1024)
        return setError(IILE_E_INVALID, std::string(who) + ": empty rectangle, or tilesize / hemiSize out of range");
    if ((int64_t)(t->x1 - t->x0) * (int64_t)(t->y1 - t->y0) > ((int64_t)1 << 26))
        return setError(IILE_E_INVALID, std::string(who) + ": task larger than 2^26 pixels");
    IisptTaskDev d{};
    d.x0 = t->x0, d.y0 = t->y0, d.x1 = t->x1, d.y1 = t->y1;
    d.tilesize = t->tilesize;
    d.hemiSize = t->hemiSize;
    d.samplerCounter = t->samplerCounter;
    d.samplerSample = t->samplerSample;
    d.rngSeed = t->rngSeed;
    iisptGrid(d, &d.gridW, &d.gridH);
    *out = d;
    return IILE_OK;
}

struct IisptBuffers {
    IisptScratch s{};
    cudaStream_t stream;
    explicit IisptBuffers(cudaStream_t st) : stream(st) {}
    cudaError_t alloc(uint32_t n, bool textured) {
        cudaError_t e = cudaMallocAsync((void **)&s.verts, n * sizeof(IisptVertex), stream);
        if (e == cudaSuccess && textured) e = cudaMallocAsync((void **)&s.texRecord, 8 * (size_t)n * sizeof(float4), stream);
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
            e = cudaMallocAsync((void **)&s.rays[k], n * sizeof(IileRay), stream);
            if (e == cudaSuccess) e = cudaMallocAsync((void **)&s.active[k], n * sizeof(uint32_t), stream);
        }
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&s.hits, n * sizeof(IileHit), stream);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&s.counter, sizeof(uint32_t), stream);
        return e;
    }
    ~IisptBuffers() {
        for (void *p : {(void *)s.verts, (void *)s.rays[0], (void *)s.rays[1], (void *)s.active[0], (void *)s.active[1],
                        (void *)s.hits, (void *)s.counter, (void *)s.texRecord})
            if (p) cudaFreeAsync(p, stream);
    }
};
}  // namespace

extern "C" {

int iile_iispt_grid(const IileIisptTask *task, int32_t *gridW, int32_t *gridH) {
    IisptTaskDev t;
    int rc = checkTask(task, "iile_iispt_grid", &t);
    if (rc != IILE_OK) return rc;
    if (gridW) *gridW = t.gridW;
    if (gridH) *gridH = t.gridH;
    return IILE_OK;
}

int iile_iispt_hemi_points(IileScene *scene, const IileIisptTask *task, float *pos, float *dir, uint8_t *valid) {
    const char *who = "iile_iispt_hemi_points";
    IisptTaskDev t;
    int rc = checkTask(task, who, &t);
    if (rc != IILE_OK) return rc;
    if (!pos || !dir || !valid) return setError(IILE_E_INVALID, std::string(who) + ": null buffer");
    if ((rc = prepare(scene, who)) != IILE_OK) return rc;
    RenderConsts rcs;
    fillConsts(scene->host.setup, &rcs);
    cudaStream_t stream = nullptr;
    DeviceScene dev;
    if ((rc = deviceViewFor(scene, stream, &dev)) != IILE_OK) return rc;
    t.stage = 0;
    const uint32_t n = (uint32_t)(t.gridW * t.gridH);
    IisptBuffers buf(stream);
    float *dPos = nullptr, *dDir = nullptr;
    uint8_t *dValid = nullptr;
    cudaError_t e = buf.alloc(n, !scene->host.materialTextures.empty());
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dPos, 3 * n * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dDir, 3 * n * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dValid, n, stream);
    uint64_t launches = 0;
    if (e == cudaSuccess) e = iisptFindVertices(dev, rcs, t, n, buf.s, stream, &launches);
    if (e == cudaSuccess) e = iisptHemiPoints(n, buf.s.verts, dPos, dDir, dValid, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pos, dPos, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dir, dDir, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(valid, dValid, n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    for (void *p : {(void *)dPos, (void *)dDir, (void *)dValid})
        if (p) cudaFreeAsync(p, stream);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, who);
}

int iile_iispt_gather_dev(IileScene *scene, const IileIisptTask *task, const float *pos, const float *dir,
                          const uint8_t *valid, const float *nnIntensity, float *out, void *streamPtr) {
    const char *who = "iile_iispt_gather_dev";
    IisptTaskDev t;
    int rc = checkTask(task, who, &t);
    if (rc != IILE_OK) return rc;
    if (!pos || !dir || !valid || !out) return setError(IILE_E_INVALID, std::string(who) + ": null buffer");
    if ((rc = prepare(scene, who)) != IILE_OK) return rc;
    RenderConsts rcs;
    fillConsts(scene->host.setup, &rcs);
    cudaStream_t stream = (cudaStream_t)streamPtr;
    DeviceScene dev;
    if ((rc = deviceViewFor(scene, stream, &dev)) != IILE_OK) return rc;
    t.stage = 1;
    const uint32_t nHemi = (uint32_t)(t.gridW * t.gridH);
    const uint32_t n = (uint32_t)((t.x1 - t.x0) * (t.y1 - t.y0));
    // per hemi point: its film (the valid ones in order) and, per film, {CameraToWorld.m, WorldToCamera.m}
    std::vector<int> hemiFilm(nHemi, -1);
    std::vector<float> cams;
    try {
        for (uint32_t h = 0; h < nHemi; ++h)
            if (valid[h]) {
                hemiFilm[h] = (int)(cams.size() / 32);
                M44 c2w, w2c, w2cInv;
                hemiCameraMatrices(pos + 3 * h, dir + 3 * h, &c2w, &w2c, &w2cInv);
                cams.resize(cams.size() + 32);
                copyM(c2w, &cams[cams.size() - 32]);
                copyM(w2c, &cams[cams.size() - 16]);
            }
    } catch (const std::exception &ex) {
        return setError(IILE_E_INVALID, std::string(who) + ": " + ex.what());
    }
    if (!cams.empty() && !nnIntensity) return setError(IILE_E_INVALID, std::string(who) + ": null nnIntensity");
    IisptBuffers buf(stream);
    float *dPos = nullptr, *dDir = nullptr, *dCams = nullptr;
    int *dFilm = nullptr;
    cudaError_t e = buf.alloc(n, !scene->host.materialTextures.empty());
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dPos, 3 * nHemi * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dDir, 3 * nHemi * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dFilm, nHemi * sizeof(int), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dCams, std::max<size_t>(cams.size(), 1) * sizeof(float), stream);
    // pageable sources: staged before the calls return
    if (e == cudaSuccess) e = cudaMemcpyAsync(dPos, pos, 3 * nHemi * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dDir, dir, 3 * nHemi * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dFilm, hemiFilm.data(), nHemi * sizeof(int), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && !cams.empty())
        e = cudaMemcpyAsync(dCams, cams.data(), cams.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    uint64_t launches = 0;
    if (e == cudaSuccess) e = iisptFindVertices(dev, rcs, t, n, buf.s, stream, &launches);
    if (e == cudaSuccess)
        e = iisptGather(dev, rcs, t, n, buf.s.verts, dPos, dDir, dFilm, dCams, nnIntensity, reinterpret_cast<float4 *>(out),
                        buf.s.texRecord, stream);
    for (void *p : {(void *)dPos, (void *)dDir, (void *)dFilm, (void *)dCams})
        if (p) cudaFreeAsync(p, stream);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, who);
}
}
