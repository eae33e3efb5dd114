// splitmethod "hlbvh" built on the GPU: BVHAccel::HLBVHBuild (src/accelerators/bvh.cpp:404-512),
// emitLBVH (bvh.cpp:514-580) and flattenBVHTree (bvh.cpp:640-658) as data-parallel kernels; the SAH over
// the (at most 4096) treelet roots, buildUpperSAH (bvh.cpp:582-638), stays on the host (hlbvhUpperTree in
// host/bvh_build.cpp).  Output = the same LinearBVHNode bytes and primitive order as the host builder and
// the reference's single-thread build (tests/test_hlbvh_device_gpu.py).
//
// Why the result is the same tree.  emitLBVH splits a Morton-sorted range at the highest bit in which its
// first and last code differ and stops at a range of fewer than maxPrimsInNode primitives or one whose codes
// are all equal: that is the binary radix tree of the sorted codes (Karras 2012) with every subtree of
// fewer than maxPrimsInNode keys, and every subtree of equal keys, collapsed into a leaf.  The radix tree
// is built with one thread per internal node (ties between equal codes broken by index, which only shapes
// the inside of subtrees that are collapsed anyway); a bottom-up pass gives each node its bounds and the
// node count of its collapsed subtree; a node's slot in the depth-first array is its treelet root's slot
// plus, for every ancestor step inside the treelet, 1 (first child) or 1 + count(first sibling) (second
// child) — at most 18 steps, so every node finds its own slot without a top-down pass.  Bounds are folded
// left to right with std::min / std::max's tie rule (the first of two equal values wins; it decides the
// sign of a zero), which is associative, so the tree-shaped fold equals the reference's serial one.
//
// Kernels: primBounds, centroidBounds (ordered-integer atomics), mortonCodes, a stable 8-bit LSD radix sort
// (count / scan / scatter with warp-match ranking), radixTree, bottomUp, treeletRoots, emitNodes.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../host/host_scene.h"

namespace iile {
namespace {

struct DBounds {
    float mn[3], mx[3];
};

// std::min(a, b) / std::max(a, b): the first argument unless the second is strictly smaller / larger
__device__ __forceinline__ float minFirst(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float maxFirst(float a, float b) { return a < b ? b : a; }

__device__ __forceinline__ DBounds uniteBounds(const DBounds &a, const DBounds &b) {
    DBounds r;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.mn[k] = minFirst(a.mn[k], b.mn[k]);
        r.mx[k] = maxFirst(a.mx[k], b.mx[k]);
    }
    return r;
}

// total order on floats as unsigned integers (-0 below +0; the sign of a zero does not reach the codes)
__device__ __forceinline__ uint32_t orderedBits(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float fromOrderedBits(uint32_t u) {
    const uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(b);
#else
    std::memcpy(&f, &b, 4);
#endif
    return f;
}

// triangle.cpp:176-186: Union(Bounds3f(p0, p1), p2)
__global__ void primBoundsKernel(const IilePrim *prims, const float *P, uint32_t n, DBounds *bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const IilePrim p = prims[i];
    if (p.type != IILE_PRIM_TRIANGLE) return;  // quadrics: bounds come from the host (scatterBoundsKernel)
    DBounds b;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a0 = P[3 * (size_t)p.v[0] + k], a1 = P[3 * (size_t)p.v[1] + k], a2 = P[3 * (size_t)p.v[2] + k];
        b.mn[k] = minFirst(minFirst(a0, a1), a2);
        b.mx[k] = maxFirst(maxFirst(a0, a1), a2);
    }
    bounds[i] = b;
}

__global__ void scatterBoundsKernel(const uint32_t *index, const DBounds *src, uint32_t n, DBounds *bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bounds[index[i]] = src[i];
}

__device__ __forceinline__ float centroidOf(const DBounds &b, int k) {
    return __fadd_rn(__fmul_rn(0.5f, b.mn[k]), __fmul_rn(0.5f, b.mx[k]));  // .5f * pMin + .5f * pMax, bvh.cpp:62
}

// cb[0..2] = min, cb[3..5] = max of the centroids, as ordered integers
__global__ void centroidBoundsKernel(const DBounds *bounds, uint32_t n, uint32_t *cb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0, 0, 0};
    if (i < n) {
        const DBounds b = bounds[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) lo[k] = hi[k] = orderedBits(centroidOf(b, k));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&cb[k], lo[k]);
            atomicMax(&cb[3 + k], hi[k]);
        }
}

__device__ __forceinline__ uint32_t leftShift3(uint32_t x) {  // bvh.cpp:118-128
    if (x == (1u << 10)) --x;
    x = (x | (x << 16)) & 0x30000ffu;
    x = (x | (x << 8)) & 0x300f00fu;
    x = (x | (x << 4)) & 0x30c30c3u;
    x = (x | (x << 2)) & 0x9249249u;
    return x;
}

struct CentroidBox {
    float mn[3], mx[3];
};

// bvh.cpp:425-433: EncodeMorton3(bounds.Offset(centroid) * 1024)
__global__ void mortonKernel(const DBounds *bounds, uint32_t n, CentroidBox cb, uint32_t *codes, uint32_t *prim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DBounds b = bounds[i];
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float o = __fsub_rn(centroidOf(b, k), cb.mn[k]);
        if (cb.mx[k] > cb.mn[k]) o = __fdiv_rn(o, __fsub_rn(cb.mx[k], cb.mn[k]));
        q[k] = __float2uint_rz(__fmul_rn(o, 1024.f));
    }
    codes[i] = (leftShift3(q[2]) << 2) | (leftShift3(q[1]) << 1) | leftShift3(q[0]);
    prim[i] = i;
}

// ---- stable LSD radix sort, 8 bits per pass (bvh.cpp:139-180 sorts 6 bits per pass; any stable sort on the
// whole code gives the same order) -------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;

__global__ void __launch_bounds__(kSortThreads) radixCountKernel(const uint32_t *keys, uint32_t n, int shift,
                                                                 uint32_t *blockHist, uint32_t numBlocks) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t idx = base + k * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    blockHist[(size_t)threadIdx.x * numBlocks + blockIdx.x] = hist[threadIdx.x];  // digit-major
}

// exclusive scan of `count` integers in place, one block of 1024 threads
__global__ void __launch_bounds__(1024) scanKernel(uint32_t *data, uint32_t count) {
    __shared__ uint32_t warpSum[32];
    const uint32_t per = (count + 1023u) / 1024u;
    const uint32_t begin = min(threadIdx.x * per, count), end = min(begin + per, count);
    uint32_t sum = 0;
    for (uint32_t i = begin; i < end; ++i) sum += data[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warpSum[lane], wi = w;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, off);
            if (lane >= off) wi += v;
        }
        warpSum[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t run = warpSum[warp] + incl - sum;
    for (uint32_t i = begin; i < end; ++i) {
        const uint32_t v = data[i];
        data[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(kSortThreads) radixScatterKernel(const uint32_t *keys, const uint32_t *vals,
                                                                   uint32_t n, int shift, const uint32_t *scanned,
                                                                   uint32_t numBlocks, uint32_t *outKeys,
                                                                   uint32_t *outVals) {
    constexpr int kWarps = kSortThreads / 32;
    constexpr int kChunks = kSortTile / kWarps / 32;  // 32-key chunks per warp
    __shared__ uint32_t warpCnt[kWarps][256];
    __shared__ uint32_t digitBase[256];
    for (int w = 0; w < kWarps; ++w) warpCnt[w][threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t segBase = blockIdx.x * kSortTile + warp * (kChunks * 32);
    uint32_t key[kChunks], rank[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const uint32_t idx = segBase + c * 32 + lane;
        const bool valid = idx < n;
        const uint32_t k = valid ? keys[idx] : 0xffffffffu;
        const uint32_t d = (k >> shift) & 255u;
        // lanes of equal digit, in lane order = array order: the stable rank inside the chunk
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0x100u + lane);
        const uint32_t r = __popc(peers & ((1u << lane) - 1u));
        const uint32_t prev = warpCnt[warp][d];
        __syncwarp();
        if (valid && r == 0) warpCnt[warp][d] = prev + __popc(peers);
        __syncwarp();
        key[c] = k;
        rank[c] = prev + r;
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = warpCnt[w][d];
            warpCnt[w][d] = run;
            run += c;
        }
        digitBase[d] = scanned[(size_t)d * numBlocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const uint32_t idx = segBase + c * 32 + lane;
        if (idx < n) {
            const uint32_t d = (key[c] >> shift) & 255u;
            const uint32_t pos = digitBase[d] + warpCnt[warp][d] + rank[c];
            outKeys[pos] = key[c];
            outVals[pos] = vals[idx];
        }
    }
}

// ---- binary radix tree over the sorted codes (Karras 2012) ------------------------------------------------
// Internal node i in [0, n-1): keys [first[i], last[i]], split position split[i]: children cover
// [first, split] and [split+1, last]; a child that covers one key is that leaf, otherwise the internal node
// numbered `split` (first child) or `split + 1` (second child).
__device__ __forceinline__ int commonPrefix(const uint32_t *codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint32_t a = codes[i], b = codes[j];
    if (a == b) return 32 + __clz(i ^ j);
    return __clz(a ^ b);
}

__global__ void radixTreeKernel(const uint32_t *codes, int n, int *first, int *last, int *split, int *internalParent,
                                int *leafParent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (commonPrefix(codes, n, i, i + 1) - commonPrefix(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int deltaMin = commonPrefix(codes, n, i, i - d);
    int lmax = 2;
    while (commonPrefix(codes, n, i, i + lmax * d) > deltaMin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (commonPrefix(codes, n, i, i + (l + t) * d) > deltaMin) l += t;
    const int j = i + l * d;
    const int deltaNode = commonPrefix(codes, n, i, j);
    int s = 0;
    for (int t = l, step; t > 1;) {
        step = (t + 1) >> 1;
        if (commonPrefix(codes, n, i, i + (s + step) * d) > deltaNode) s += step;
        t = step;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    first[i] = lo;
    last[i] = hi;
    split[i] = gamma;
    if (lo == gamma) leafParent[gamma] = i;
    else internalParent[gamma] = i;
    if (hi == gamma + 1) leafParent[gamma + 1] = i;
    else internalParent[gamma + 1] = i;
    if (i == 0) internalParent[0] = -1;
}

// prefix length of a node's range on the 32-bit code words (two leading zero bits + 30 code bits); 32 when
// every code of the range is the same
__device__ __forceinline__ int rangePrefix(const uint32_t *codes, int f, int l) {
    const uint32_t a = codes[f], b = codes[l];
    return a == b ? 32 : __clz(a ^ b);
}
constexpr int kTreeletPrefix = 2 + 12;  // ranges with at least this prefix lie inside one treelet (bvh.cpp:442)

// Each leaf climbs; the second thread to reach a node finds both children finished.
__global__ void bottomUpKernel(const uint32_t *codes, const uint32_t *prim, const DBounds *primBounds, int n,
                               int maxPrims, const int *first, const int *last, const int *split,
                               const int *internalParent, const int *leafParent, uint32_t *visits,
                               DBounds *nodeBounds, int *nodeCount, int *maxLeaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cur = leafParent[i];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(&visits[cur], 1u) == 0u) return;
        __threadfence();
        const int f = first[cur], l = last[cur], g = split[cur];
        DBounds a, b;
        int ca, cb;
        if (f == g) {
            a = primBounds[prim[g]];
            ca = 1;
        } else {
            const float *src = reinterpret_cast<const float *>(&nodeBounds[g]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a.mn[k] = __ldcg(src + k);
                a.mx[k] = __ldcg(src + 3 + k);
            }
            ca = __ldcg(&nodeCount[g]);
        }
        if (l == g + 1) {
            b = primBounds[prim[g + 1]];
            cb = 1;
        } else {
            const float *src = reinterpret_cast<const float *>(&nodeBounds[g + 1]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                b.mn[k] = __ldcg(src + k);
                b.mx[k] = __ldcg(src + 3 + k);
            }
            cb = __ldcg(&nodeCount[g + 1]);
        }
        const int prefix = rangePrefix(codes, f, l), count = l - f + 1;
        const bool leaf = count < maxPrims || prefix >= 32;  // bvh.cpp:519: bit == -1 || nPrimitives < maxPrimsInNode
        if (leaf && prefix >= kTreeletPrefix) atomicMax(maxLeaf, count);
        nodeBounds[cur] = uniteBounds(a, b);
        nodeCount[cur] = leaf ? 1 : 1 + ca + cb;
        cur = internalParent[cur];
    }
}

// One entry per value of the codes' top 12 bits.  count: nodes of the treelet, 0 = no primitive carries that prefix.
// ref: the treelet's root — an internal node of the radix tree, or kSingleTreelet for a treelet of one primitive
// (a leaf), or kNoTreeletRoot if no node spans the treelet's range (an internal error, reported by the host).  The
// first version encoded a single primitive at sorted position s as ~s, which collides with the two markers for
// s = 0 and s = 1: a lone primitive in the first treelet (a far-away sphere among triangles) was dropped, or reported
// as an error (found by tools/fuzz_bvh.py).
constexpr int kSingleTreelet = -3, kNoTreeletRoot = -2;
struct TreeletRoot {
    int ref;
    int count;
    DBounds bounds;
};

__global__ void treeletRangeKernel(const uint32_t *codes, int n, int *start, int *end) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = codes[i] >> 18;
    if (i == 0 || (codes[i - 1] >> 18) != v) start[v] = i;
    if (i == n - 1 || (codes[i + 1] >> 18) != v) end[v] = i;
}

__global__ void treeletRootKernel(const uint32_t *prim, const DBounds *primBounds, const int *start, const int *end,
                                  const int *first, const int *last, const DBounds *nodeBounds, const int *nodeCount,
                                  TreeletRoot *roots) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 4096) return;
    TreeletRoot r;
    r.ref = -1;
    r.count = 0;
    r.bounds = DBounds{};
    const int s = start[v], e = end[v];
    if (s >= 0) {
        if (s == e) {
            r.ref = kSingleTreelet;
            r.count = 1;
            r.bounds = primBounds[prim[s]];
        } else {
            // the internal node of a range is numbered by one of the range's ends
            const int node = (first[s] == s && last[s] == e) ? s : e;
            r.ref = (first[node] == s && last[node] == e) ? node : kNoTreeletRoot;  // cannot happen (checked on the host)
            r.count = nodeCount[node];
            r.bounds = nodeBounds[node];
        }
    }
    roots[v] = r;
}

// Write every node of the collapsed tree into its depth-first slot.  Threads [0, n-1) are the radix tree's
// internal nodes, threads [n-1, 2n-1) its leaves.
__global__ void emitNodesKernel(const uint32_t *codes, const uint32_t *prim, const DBounds *primBounds, int n,
                                int maxPrims, const int *first, const int *last, const int *split,
                                const int *internalParent, const int *leafParent, const DBounds *nodeBounds,
                                const int *nodeCount, const int *rootSlot, IileBVHNode *out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n - 1) return;
    const bool isLeafThread = t >= n - 1;
    const int self = isLeafThread ? t - (n - 1) : t;
    int f, l, parent;
    if (isLeafThread) {
        f = l = self;
        parent = leafParent[self];
    } else {
        f = first[self];
        l = last[self];
        parent = internalParent[self];
        if (rangePrefix(codes, f, l) < kTreeletPrefix) return;  // spans treelets: the upper SAH tree's business
    }
    // part of the collapsed tree: a treelet root, or the child of a node that is interior there
    if (parent >= 0) {
        const int pf = first[parent], pl = last[parent];
        const int pp = rangePrefix(codes, pf, pl);
        if (pp >= kTreeletPrefix && (pp >= 32 || pl - pf + 1 < maxPrims)) return;  // inside a collapsed leaf
    }
    int slot = 0;
    {
        int cur = self;
        bool curIsLeaf = isLeafThread;
        int p = parent;
        while (p >= 0) {
            const int pf = first[p], pl = last[p], g = split[p];
            if (rangePrefix(codes, pf, pl) < kTreeletPrefix) break;  // cur is its treelet's root
            const bool firstChild = cur == g && (curIsLeaf == (pf == g));
            slot += firstChild ? 1 : 1 + (pf == g ? 1 : nodeCount[g]);
            cur = p;
            curIsLeaf = false;
            p = internalParent[p];
        }
    }
    slot += rootSlot[codes[f] >> 18];
    IileBVHNode node;
    memset(&node, 0, sizeof(node));
    DBounds b;
    const int prefix = isLeafThread ? 32 : rangePrefix(codes, f, l);
    const int count = l - f + 1;
    if (isLeafThread) b = primBounds[prim[self]];
    else b = nodeBounds[self];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        node.bmin[k] = b.mn[k];
        node.bmax[k] = b.mx[k];
    }
    if (count < maxPrims || prefix >= 32) {  // leaf: primitives [f, l] of the sorted order
        node.offset = f;
        node.nPrims = (uint16_t)count;
    } else {
        const int g = split[self];
        node.axis = (uint8_t)((31 - prefix) % 3);  // the splitting bit, bvh.cpp:553
        node.offset = slot + 1 + (f == g ? 1 : nodeCount[g]);
    }
    out[slot] = node;
}

// One device allocation carved into the build's buffers (a single cudaMalloc / cudaFree: freeing twenty
// large blocks one by one costs more than the build in a process that holds other allocations).
struct DeviceArena {
    char *base = nullptr;
    size_t size = 0, used = 0;
    ~DeviceArena() {
        if (base) cudaFree(base);
    }
    static size_t aligned(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
    bool reserve(size_t bytes) {
        size = bytes;
        return cudaMalloc((void **)&base, bytes) == cudaSuccess;
    }
    template <typename T>
    T *get(size_t count) {
        const size_t bytes = aligned(std::max<size_t>(count, 1) * sizeof(T));
        if (used + bytes > size) return nullptr;
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
};

#define HL_CHECK(expr)                                                                  \
    do {                                                                                \
        const cudaError_t err__ = (expr);                                               \
        if (err__ != cudaSuccess) {                                                     \
            *why = std::string(#expr) + ": " + cudaGetErrorString(err__);               \
            return false;                                                               \
        }                                                                               \
    } while (0)

bool deviceHlbvh(HostScene *scene, std::string *why) {
    HostScene &s = *scene;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        *why = "no CUDA device";
        return false;
    }
    const size_t nz = s.prims.size();
    if (nz < 2 || nz >= (size_t)1 << 30) {
        *why = "primitive count out of range";
        return false;
    }
    const int n = (int)nz;
    const int maxPrims = std::min(255, s.setup.maxNodePrims);
    const bool trace = std::getenv("IILE_BUILD_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "hlbvh(device): %s %.3f s\n", what, std::chrono::duration<double>(now - t0).count());
        t0 = now;
    };

    const uint32_t sortBlocks = (uint32_t)((nz + kSortTile - 1) / kSortTile);
    size_t nQuadrics = 0;
    for (const IilePrim &p : s.prims) nQuadrics += p.type != IILE_PRIM_TRIANGLE;
    DeviceArena mem;
    {
        auto A = [](size_t bytes) { return DeviceArena::aligned(std::max<size_t>(bytes, 1)); };
        const size_t total = A(nz * sizeof(IilePrim)) + A(s.P.size() * sizeof(float)) + 2 * A(nz * sizeof(DBounds)) +
                             A(6 * 4) + 4 * A(nz * 4) + A((size_t)256 * sortBlocks * 4) + 7 * A(nz * 4) + A(4) +
                             3 * A(4096 * 4) + A(4096 * sizeof(TreeletRoot)) + A(nQuadrics * 4) +
                             A(nQuadrics * sizeof(DBounds)) + A((2 * nz + 8192) * sizeof(IileBVHNode));
        if (!mem.reserve(total)) {
            cudaGetLastError();
            *why = "out of device memory";
            return false;
        }
    }
    IilePrim *dPrims = mem.get<IilePrim>(nz);
    float *dP = mem.get<float>(s.P.size());
    DBounds *dBounds = mem.get<DBounds>(nz);
    uint32_t *dCb = mem.get<uint32_t>(6);
    uint32_t *keysA = mem.get<uint32_t>(nz), *keysB = mem.get<uint32_t>(nz);
    uint32_t *valsA = mem.get<uint32_t>(nz), *valsB = mem.get<uint32_t>(nz);
    uint32_t *dHist = mem.get<uint32_t>((size_t)256 * sortBlocks);
    int *dFirst = mem.get<int>(nz), *dLast = mem.get<int>(nz), *dSplit = mem.get<int>(nz);
    int *dIntParent = mem.get<int>(nz), *dLeafParent = mem.get<int>(nz);
    uint32_t *dVisits = mem.get<uint32_t>(nz);
    DBounds *dNodeBounds = mem.get<DBounds>(nz);
    int *dNodeCount = mem.get<int>(nz);
    int *dMaxLeaf = mem.get<int>(1);
    int *dStart = mem.get<int>(4096), *dEnd = mem.get<int>(4096), *dRootSlot = mem.get<int>(4096);
    TreeletRoot *dRoots = mem.get<TreeletRoot>(4096);
    if (!dPrims || !dP || !dBounds || !dCb || !keysA || !keysB || !valsA || !valsB || !dHist || !dFirst || !dLast ||
        !dSplit || !dIntParent || !dLeafParent || !dVisits || !dNodeBounds || !dNodeCount || !dMaxLeaf || !dStart ||
        !dEnd || !dRootSlot || !dRoots) {
        cudaGetLastError();
        *why = "out of device memory";
        return false;
    }
    HL_CHECK(cudaMemcpy(dPrims, s.prims.data(), nz * sizeof(IilePrim), cudaMemcpyHostToDevice));
    HL_CHECK(cudaMemcpy(dP, s.P.data(), s.P.size() * sizeof(float), cudaMemcpyHostToDevice));
    lap("upload");

    const int threads = 256;
    const uint32_t blocksN = (uint32_t)((nz + threads - 1) / threads);
    primBoundsKernel<<<blocksN, threads>>>(dPrims, dP, (uint32_t)n, dBounds);
    {  // quadrics: their world bounds need the object-to-world transform (sphere.cpp:44-47); few of them
        std::vector<uint32_t> index;
        std::vector<DBounds> bounds;
        for (int i = 0; i < n; ++i)
            if (s.prims[i].type != IILE_PRIM_TRIANGLE) {
                const Bounds3 wb = primWorldBound(s, s.prims[i]);
                index.push_back((uint32_t)i);
                bounds.push_back(DBounds{{wb.pMin.x, wb.pMin.y, wb.pMin.z}, {wb.pMax.x, wb.pMax.y, wb.pMax.z}});
            }
        if (!index.empty()) {
            uint32_t *dIndex = mem.get<uint32_t>(index.size());
            DBounds *dSrc = mem.get<DBounds>(index.size());
            if (!dIndex || !dSrc) {
                cudaGetLastError();
                *why = "out of device memory";
                return false;
            }
            HL_CHECK(cudaMemcpy(dIndex, index.data(), index.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            HL_CHECK(cudaMemcpy(dSrc, bounds.data(), bounds.size() * sizeof(DBounds), cudaMemcpyHostToDevice));
            scatterBoundsKernel<<<(uint32_t)((index.size() + threads - 1) / threads), threads>>>(
                dIndex, dSrc, (uint32_t)index.size(), dBounds);
        }
    }
    {
        const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        HL_CHECK(cudaMemcpy(dCb, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    centroidBoundsKernel<<<blocksN, threads>>>(dBounds, (uint32_t)n, dCb);
    CentroidBox cb;
    {
        uint32_t raw[6];
        HL_CHECK(cudaMemcpy(raw, dCb, sizeof(raw), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; ++k) {
            cb.mn[k] = fromOrderedBits(raw[k]);
            cb.mx[k] = fromOrderedBits(raw[3 + k]);
        }
    }
    mortonKernel<<<blocksN, threads>>>(dBounds, (uint32_t)n, cb, keysA, valsA);
    lap("bounds + codes");

    for (int pass = 0; pass < 4; ++pass) {
        const uint32_t *inK = (pass & 1) ? keysB : keysA, *inV = (pass & 1) ? valsB : valsA;
        uint32_t *outK = (pass & 1) ? keysA : keysB, *outV = (pass & 1) ? valsA : valsB;
        radixCountKernel<<<sortBlocks, kSortThreads>>>(inK, (uint32_t)n, 8 * pass, dHist, sortBlocks);
        scanKernel<<<1, 1024>>>(dHist, 256u * sortBlocks);
        radixScatterKernel<<<sortBlocks, kSortThreads>>>(inK, inV, (uint32_t)n, 8 * pass, dHist, sortBlocks, outK, outV);
    }
    const uint32_t *codes = keysA, *prim = valsA;  // four passes: back in A
    HL_CHECK(cudaGetLastError());
    lap("radix sort");

    HL_CHECK(cudaMemset(dVisits, 0, nz * sizeof(uint32_t)));
    HL_CHECK(cudaMemset(dMaxLeaf, 0, sizeof(int)));
    HL_CHECK(cudaMemset(dStart, 0xff, 4096 * sizeof(int)));
    HL_CHECK(cudaMemset(dEnd, 0xff, 4096 * sizeof(int)));
    radixTreeKernel<<<blocksN, threads>>>(codes, n, dFirst, dLast, dSplit, dIntParent, dLeafParent);
    bottomUpKernel<<<blocksN, threads>>>(codes, prim, dBounds, n, maxPrims, dFirst, dLast, dSplit, dIntParent,
                                         dLeafParent, dVisits, dNodeBounds, dNodeCount, dMaxLeaf);
    treeletRangeKernel<<<blocksN, threads>>>(codes, n, dStart, dEnd);
    treeletRootKernel<<<4096 / threads, threads>>>(prim, dBounds, dStart, dEnd, dFirst, dLast, dNodeBounds, dNodeCount,
                                                   dRoots);
    std::vector<TreeletRoot> roots(4096);
    HL_CHECK(cudaMemcpy(roots.data(), dRoots, 4096 * sizeof(TreeletRoot), cudaMemcpyDeviceToHost));
    int maxLeaf = 0;
    HL_CHECK(cudaMemcpy(&maxLeaf, dMaxLeaf, sizeof(int), cudaMemcpyDeviceToHost));
    lap("radix tree + bounds");
    if (maxLeaf >= 65536) {
        *why = "a treelet leaf holds 65536 primitives or more";  // the host build reports the error
        return false;
    }

    // buildUpperSAH over the treelets, in the order of their Morton prefix (bvh.cpp:440-460 walks the sorted
    // array), then the slot of every upper node and treelet root by the depth-first rule of flattenBVHTree
    std::vector<int> prefixOf;
    std::vector<Bounds3> rootBounds;
    std::vector<int> rootCounts;
    for (int v = 0; v < 4096; ++v) {
        const TreeletRoot &r = roots[v];
        if (r.count == 0) continue;
        if (r.ref == kNoTreeletRoot) {
            *why = "internal error: treelet root not found in the radix tree";
            return false;
        }
        prefixOf.push_back(v);
        rootBounds.push_back(Bounds3(V3(r.bounds.mn[0], r.bounds.mn[1], r.bounds.mn[2]),
                                     V3(r.bounds.mx[0], r.bounds.mx[1], r.bounds.mx[2])));
        rootCounts.push_back(r.count);
    }
    const int nTreelets = (int)prefixOf.size();
    std::vector<HlbvhUpperNode> upper;
    int upperRoot = 0;
    try {
        hlbvhUpperTree(rootBounds, rootCounts, &upper, &upperRoot);
    } catch (const std::exception &ex) {
        *why = ex.what();  // the host build raises the same error
        return false;
    }
    const int total = upper[upperRoot].count;
    std::vector<int> rootSlot(4096, -1);
    struct UpperSlot {
        int node, slot;
    };
    std::vector<UpperSlot> upperSlots;  // interior nodes of the upper tree
    {
        std::vector<UpperSlot> stack{{upperRoot, 0}};
        while (!stack.empty()) {
            const UpperSlot it = stack.back();
            stack.pop_back();
            if (it.node < nTreelets) {
                rootSlot[prefixOf[it.node]] = it.slot;
                continue;
            }
            upperSlots.push_back(it);
            const HlbvhUpperNode &u = upper[it.node];
            stack.push_back({u.child[1], it.slot + 1 + upper[u.child[0]].count});
            stack.push_back({u.child[0], it.slot + 1});
        }
    }
    HL_CHECK(cudaMemcpy(dRootSlot, rootSlot.data(), 4096 * sizeof(int), cudaMemcpyHostToDevice));
    IileBVHNode *dOut = mem.get<IileBVHNode>((size_t)total);
    if (!dOut) {
        cudaGetLastError();
        *why = "out of device memory";
        return false;
    }
    {
        const size_t threadsTotal = 2 * nz - 1;
        emitNodesKernel<<<(uint32_t)((threadsTotal + threads - 1) / threads), threads>>>(
            codes, prim, dBounds, n, maxPrims, dFirst, dLast, dSplit, dIntParent, dLeafParent, dNodeBounds, dNodeCount,
            dRootSlot, dOut);
    }
    HL_CHECK(cudaGetLastError());
    lap("upper SAH + emit");

    s.nodes.resize((size_t)total);
    s.primOrder.resize(nz);
    HL_CHECK(cudaMemcpy(s.nodes.data(), dOut, (size_t)total * sizeof(IileBVHNode), cudaMemcpyDeviceToHost));
    HL_CHECK(cudaMemcpy(s.primOrder.data(), prim, nz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (const UpperSlot &it : upperSlots) {
        const HlbvhUpperNode &u = upper[it.node];
        IileBVHNode &ln = s.nodes[it.slot];
        std::memset(&ln, 0, sizeof(ln));
        for (int k = 0; k < 3; ++k) {
            ln.bmin[k] = u.bounds.pMin[k];
            ln.bmax[k] = u.bounds.pMax[k];
        }
        ln.axis = (uint8_t)u.axis;
        ln.offset = it.slot + 1 + upper[u.child[0]].count;
    }
    lap("download");
    return true;
}

const bool g_installed = (g_deviceHlbvh = &deviceHlbvh, true);

}  // namespace
}  // namespace iile
