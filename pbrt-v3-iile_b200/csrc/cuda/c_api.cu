// C ABI (include/iile_b200.h): scene lifetime, upload, and the ray-cast entry
// points.  The render entry points live in render_api.cu.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <exception>
#include <memory>
#include <thread>
#include <limits>
#include <string>
#include <vector>

#include "cast_kernels.h"
#include "scene_handle.h"

using namespace iile;

namespace iile {

thread_local std::string g_lastError;

int setError(int code, const std::string &msg) {
    g_lastError = msg;
    return code;
}

int cudaFail(cudaError_t e, const char *what) {
    return setError(IILE_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

template <typename T>
static cudaError_t uploadVec(const std::vector<T> &v, T **dst) {
    *dst = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void **)dst, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

void IileSceneFree(IileScene *s) {
    if (!s) return;
    freeCastStaging(s);
    for (void *p : s->deviceAllocs) cudaFree(p);
    s->deviceAllocs.clear();
    s->counterStreams.clear();
    s->uploaded = false;
}

int deviceViewFor(IileScene *s, cudaStream_t stream, DeviceScene *out) {
    std::lock_guard<std::mutex> lock(s->counterMutex);
    size_t k = 0;
    while (k < s->counterStreams.size() && s->counterStreams[k] != stream) ++k;
    if (k == s->counterStreams.size()) {
        if (k >= (size_t)kCounterPairs)
            return setError(IILE_E_UNSUPPORTED, "more than 1024 distinct CUDA streams have cast rays against this scene");
        s->counterStreams.push_back(stream);
    }
    *out = s->dev;
    out->castCounters = s->dev.castCounters + 2 * k;
    return IILE_OK;
}

// Fuse two levels of the binary tree into 128-byte 4-wide nodes (layout: traverse.cuh).
// Returns the deepest wide level through *depth.
struct WideNode {
    float box[6][4];  // minx maxx miny maxy minz maxz, one lane per slot
    uint32_t ref[4];
    uint32_t meta, pad[3];
};
static_assert(sizeof(WideNode) == 128, "wide node must be one 128-byte line");

// A leaf of more than kLeafChunk primitives (the reference allows 65,535: uint16 nPrimitives,
// bvh.cpp:95-104 — e.g. the coincident-centroid leaf of recursiveBuild, bvh.cpp:262-277, or an HLBVH
// treelet whose primitives share a Morton cell, bvh.cpp:487-505) does not fit the 8-bit count of a
// leaf reference.  It becomes a chain of CHUNK nodes: a wide node whose slots are leaf references to
// consecutive runs of <= 255 primitives (the fourth slot continues the chain when more remain).
// Chunk slots carry the box [-FLT_MAX, +FLT_MAX], which every ray passes with entry distance -inf, so
// a popped chunk is never culled against tMax — the reference tests the leaf's box once (that test is
// the parent slot's, with the leaf's real box) and then every primitive of the leaf.  The three axis
// fields of a chunk node's meta are 3: bit 3 of a ray's negMask is always clear, so the slots are
// visited 0,1,2,3 for every ray — array order, as bvh.cpp:674-679 does.
constexpr uint32_t kLeafChunk = 255;

// Run fn(begin, end) over [0, n) cut into contiguous pieces on the host's cores.
template <typename F>
static void hostParallelFor(size_t n, size_t minPerThread, F fn) {
    const char *env = std::getenv("IILE_BUILD_THREADS");
    size_t threads = env ? (size_t)std::max(1, std::atoi(env)) : (size_t)std::thread::hardware_concurrency();
    threads = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(threads, 64), n / std::max<size_t>(minPerThread, 1)));
    if (threads <= 1) {
        fn((size_t)0, n);
        return;
    }
    // an exception inside a worker (std::bad_alloc while a subtree's nodes grow) is carried to the caller
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> failed(threads);
    for (size_t t = 0; t < threads; ++t)
        th.emplace_back([=, &failed] {
            try {
                fn(n * t / threads, n * (t + 1) / threads);
            } catch (...) {
                failed[t] = std::current_exception();
            }
        });
    for (auto &x : th) x.join();
    for (const std::exception_ptr &e : failed)
        if (e) std::rethrow_exception(e);
}

// Builds the wide nodes of one binary subtree into `out` (references are indices into `out`).  With
// `deferred` set, a child whose binary subtree has at most `deferBelow` nodes is not expanded but recorded
// as a task: the slot that refers to it is patched once the task's nodes have a place in the final array.
struct WideBuilder {
    const std::vector<IileBVHNode> &nodes;
    std::vector<WideNode> out;
    int depth = 0;        // deepest wide level reached (root level counted from the whole tree's root)
    int chunkLevels = 0;  // deepest chain of chunk nodes below any leaf
    struct Task {
        int node, end;    // binary subtree [node, end)
        int wide, slot;   // the reference to patch
        int level;
    };
    explicit WideBuilder(const std::vector<IileBVHNode> &n) : nodes(n) {}

    static WideNode emptyNode() {
        WideNode w{};
        for (int slot = 0; slot < 4; ++slot)
            for (int a = 0; a < 3; ++a) {
                w.box[2 * a][slot] = std::numeric_limits<float>::max();
                w.box[2 * a + 1][slot] = -std::numeric_limits<float>::max();
            }
        return w;
    }
    static void setBox(WideNode &w, int slot, const IileBVHNode &n) {
        for (int a = 0; a < 3; ++a) {
            w.box[2 * a][slot] = n.bmin[a];
            w.box[2 * a + 1][slot] = n.bmax[a];
        }
    }
    // reference to `count` primitives starting at leaf slot `first`: a plain leaf reference, or the
    // index of the first node of a chunk chain appended to out
    uint32_t leafRef(uint32_t first, uint32_t count) {
        if (count <= kLeafChunk) return 0x80000000u | (count << 23) | first;
        const uint32_t head = (uint32_t)out.size();
        int levels = 0;
        while (count > 0) {
            ++levels;
            const size_t at = out.size();
            out.push_back(emptyNode());
            WideNode w = emptyNode();
            w.meta = 3u | (3u << 2) | (3u << 4);
            for (int slot = 0; slot < 4 && count > 0; ++slot) {
                for (int a = 0; a < 3; ++a) {
                    w.box[2 * a][slot] = -std::numeric_limits<float>::max();
                    w.box[2 * a + 1][slot] = std::numeric_limits<float>::max();
                }
                w.meta |= 0x100u << slot;
                if (slot == 3 && count > kLeafChunk) {
                    w.ref[slot] = (uint32_t)out.size();  // the chain continues in the node appended next
                    break;
                }
                const uint32_t take = std::min(count, kLeafChunk);
                w.ref[slot] = 0x80000000u | (take << 23) | first;
                first += take;
                count -= take;
            }
            out[at] = w;
        }
        chunkLevels = std::max(chunkLevels, levels);
        return head;
    }

    void run(int rootNode, int rootEnd, int rootLevel, int deferBelow, std::vector<Task> *deferred) {
        struct Todo { int node, end, wide, level; };
        std::vector<Todo> todo;
        out.push_back(WideNode{});
        todo.push_back({rootNode, rootEnd, 0, rootLevel});
        while (!todo.empty()) {
            const Todo t = todo.back();
            todo.pop_back();
            depth = std::max(depth, t.level);
            const IileBVHNode &N = nodes[t.node];
            WideNode w = emptyNode();
            w.meta = N.axis & 3u;
            const int group[2] = {t.node + 1, N.offset}, groupEnd[2] = {N.offset, t.end};
            for (int g = 0; g < 2; ++g) {
                const IileBVHNode &X = nodes[group[g]];
                int child[2], childEnd[2], nChild;
                if (X.nPrims == 0) {
                    child[0] = group[g] + 1;
                    childEnd[0] = X.offset;
                    child[1] = X.offset;
                    childEnd[1] = groupEnd[g];
                    nChild = 2;
                    w.meta |= (X.axis & 3u) << (2 + 2 * g);
                } else {
                    child[0] = group[g];
                    childEnd[0] = groupEnd[g];
                    nChild = 1;
                }
                for (int j = 0; j < nChild; ++j) {
                    const int slot = 2 * g + j;
                    const IileBVHNode &C = nodes[child[j]];
                    setBox(w, slot, C);
                    w.meta |= 0x100u << slot;
                    if (C.nPrims > 0) w.ref[slot] = leafRef((uint32_t)C.offset, C.nPrims);
                    else if (deferred && childEnd[j] - child[j] <= deferBelow) {
                        w.ref[slot] = 0;
                        deferred->push_back({child[j], childEnd[j], t.wide, slot, t.level + 1});
                    } else {
                        w.ref[slot] = (uint32_t)out.size();
                        todo.push_back({child[j], childEnd[j], (int)out.size(), t.level + 1});
                        out.push_back(WideNode{});
                    }
                }
            }
            out[t.wide] = w;
        }
    }
};

static bool wideInterior(const WideNode &w, int slot) {
    return (w.meta & (0x100u << slot)) && !(w.ref[slot] & 0x80000000u);
}

// The whole tree.  Large trees are cut below the top into subtrees that the host's cores build
// concurrently, each into its own block of the final array (top part first, then the blocks in the order
// the top pass met them); the numbering only decides where a node lives, not what the kernels compute.
static void buildWideNodes(const std::vector<IileBVHNode> &nodes, std::vector<WideNode> *out, int *depth) {
    out->clear();
    *depth = 0;
    if (nodes.empty()) return;
    WideBuilder top(nodes);
    if (nodes[0].nPrims > 0) {  // the whole scene is one leaf
        top.out.push_back(WideBuilder::emptyNode());
        WideNode w = WideBuilder::emptyNode();
        WideBuilder::setBox(w, 0, nodes[0]);
        w.ref[0] = top.leafRef((uint32_t)nodes[0].offset, nodes[0].nPrims);
        w.meta = 0x100u;
        top.out[0] = w;
        *depth = 1 + top.chunkLevels;
        out->swap(top.out);
        return;
    }
    const int total = (int)nodes.size();
    if (total < (1 << 18)) {
        top.run(0, total, 1, 0, nullptr);
        *depth = top.depth + top.chunkLevels;
        out->swap(top.out);
        return;
    }
    std::vector<WideBuilder::Task> tasks;
    top.run(0, total, 1, std::max(total / 256, 1 << 12), &tasks);
    std::vector<std::unique_ptr<WideBuilder>> parts(tasks.size());
    {
        std::atomic<size_t> next{0};
        hostParallelFor(tasks.size(), 1, [&](size_t, size_t) {
            for (size_t k; (k = next.fetch_add(1)) < tasks.size();) {
                parts[k].reset(new WideBuilder(nodes));
                parts[k]->run(tasks[k].node, tasks[k].end, tasks[k].level, 0, nullptr);
            }
        });
    }
    std::vector<size_t> base(tasks.size() + 1);
    base[0] = top.out.size();
    int deepest = top.depth, chunks = top.chunkLevels;
    for (size_t k = 0; k < tasks.size(); ++k) {
        base[k + 1] = base[k] + parts[k]->out.size();
        deepest = std::max(deepest, parts[k]->depth);
        chunks = std::max(chunks, parts[k]->chunkLevels);
        top.out[tasks[k].wide].ref[tasks[k].slot] = (uint32_t)base[k];
    }
    *depth = deepest + chunks;
    out->swap(top.out);
    out->resize(base.back());
    {
        std::atomic<size_t> next{0};
        hostParallelFor(tasks.size(), 1, [&](size_t, size_t) {
            for (size_t k; (k = next.fetch_add(1)) < tasks.size();) {
                const uint32_t shift = (uint32_t)base[k];
                WideNode *dst = out->data() + base[k];
                const std::vector<WideNode> &src = parts[k]->out;
                for (size_t i = 0; i < src.size(); ++i) {
                    WideNode w = src[i];
                    for (int slot = 0; slot < 4; ++slot)
                        if (wideInterior(w, slot)) w.ref[slot] += shift;
                    dst[i] = w;
                }
                parts[k].reset();
            }
        });
    }
}

// The top of the wide tree, breadth first, takes the first kTopStaged slots of the array (the breadth-first
// numbering is what IILE_TOP_SMEM stages in shared memory, traverse.cuh).  The nodes that sat in those slots
// move to the places the breadth-first nodes leave; every other node stays where it is.  A node's parent
// has a smaller index than the node (slots are handed out when the parent is laid out), so every reference
// that has to change sits in one of the nodes that move.
static void relabelTopBreadthFirst(std::vector<WideNode> *nodes, size_t limit) {
    const size_t n = nodes->size();
    if (n <= 1) return;
    std::vector<uint32_t> bfs;
    bfs.push_back(0);
    for (size_t head = 0; head < bfs.size() && bfs.size() < limit; ++head)
        for (int slot = 0; slot < 4 && bfs.size() < limit; ++slot)
            if (wideInterior((*nodes)[bfs[head]], slot)) bfs.push_back((*nodes)[bfs[head]].ref[slot]);
    const size_t m = bfs.size();
    std::vector<std::pair<uint32_t, uint32_t>> moved;  // old index -> new index
    std::vector<char> inBfs(m, 0);                     // which of the slots [0, m) hold a breadth-first node already
    std::vector<uint32_t> vacated;                     // breadth-first nodes that lived outside [0, m)
    for (size_t k = 0; k < m; ++k) {
        moved.push_back({bfs[k], (uint32_t)k});
        if (bfs[k] < m) inBfs[bfs[k]] = 1;
        else vacated.push_back(bfs[k]);
    }
    std::sort(vacated.begin(), vacated.end());
    size_t v = 0;
    for (size_t i = 0; i < m; ++i)
        if (!inBfs[i]) moved.push_back({(uint32_t)i, vacated[v++]});
    std::sort(moved.begin(), moved.end());
    auto newIndex = [&](uint32_t old) {
        auto it = std::lower_bound(moved.begin(), moved.end(), std::make_pair(old, 0u));
        return (it != moved.end() && it->first == old) ? it->second : old;
    };
    std::vector<WideNode> copies;
    for (const auto &mv : moved) {
        WideNode w = (*nodes)[mv.first];
        for (int slot = 0; slot < 4; ++slot)
            if (wideInterior(w, slot)) w.ref[slot] = newIndex(w.ref[slot]);
        copies.push_back(w);
    }
    for (size_t k = 0; k < moved.size(); ++k) (*nodes)[moved[k].second] = copies[k];
}

static int uploadSceneImpl(IileScene *s);

// Lay the scene out in HBM (see device_scene.h).  Host-side allocation failures while the device layout is
// built (wide nodes, leaf slots) come back as an error code, not as an exception through the C ABI.
int uploadScene(IileScene *s) {
    try {
        return uploadSceneImpl(s);
    } catch (const std::exception &ex) {
        IileSceneFree(s);
        return setError(IILE_E_INVALID, std::string("scene upload: ") + ex.what());
    }
}

static int uploadSceneImpl(IileScene *s) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return setError(IILE_E_CUDA, std::string("no CUDA device available: ") +
                                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    const HostScene &h = s->host;
    DeviceScene &d = s->dev;
    std::memset(&d, 0, sizeof(d));
    if (h.prims.size() >= (1u << 23) - 1)  // before anything is allocated: nothing to free on this path
        return setError(IILE_E_UNSUPPORTED, "scenes beyond 8,388,606 primitives exceed the 23-bit leaf "
                                            "offset of the wide-node encoding");
    const bool trace = std::getenv("IILE_BUILD_TRACE") != nullptr;
    auto tPhase = std::chrono::steady_clock::now();
    auto phase = [&](const char *what) {
        const auto now = std::chrono::steady_clock::now();
        if (trace) std::fprintf(stderr, "upload: %s %.3f s\n", what, std::chrono::duration<double>(now - tPhase).count());
        tPhase = now;
    };
    // leaf-order primitive slots: pre-gathered triangle vertices + ids in the .w lanes
    std::vector<float4> leaf(3 * h.prims.size());
    hostParallelFor(h.primOrder.size(), 1 << 16, [&](size_t slotBegin, size_t slotEnd) {
    for (size_t slot = slotBegin; slot < slotEnd; ++slot) {
        uint32_t pi = h.primOrder[slot];
        const IilePrim &p = h.prims[pi];
        float4 a = {0, 0, 0, 0}, b = a, c = a;
        uint32_t kind = p.flags & 0xffff;
        // shading class of the primitive's material (wavefront.h: kClassMatte.. = type + 1)
        // quadrics always go to the generic kernel (kClassOther): the per-material kernels inline the
        // triangle surface only
        // (so do materials with textures: texture.cuh's resolveTextured runs in the generic kernel)
        kind |= (uint32_t)((p.type == IILE_PRIM_TRIANGLE && p.material >= 0 && p.material < (int)h.materials.size() &&
                            !h.materials[p.material].texSet)
                               ? std::min(h.materials[p.material].type + 1, 5) : 5) << 16;
        int aux = p.material;
        if (p.type == IILE_PRIM_TRIANGLE) {
            V3 p0 = h.vertex(p.v[0]), p1 = h.vertex(p.v[1]), p2 = h.vertex(p.v[2]);
            a = {p0.x, p0.y, p0.z, 0};
            b = {p1.x, p1.y, p1.z, 0};
            c = {p2.x, p2.y, p2.z, 0};
        } else {
            kind |= kLeafSphere;
            aux = p.v[0];
        }
        std::memcpy(&a.w, &pi, 4);
        std::memcpy(&b.w, &kind, 4);
        std::memcpy(&c.w, &aux, 4);
        leaf[3 * slot] = a;
        leaf[3 * slot + 1] = b;
        leaf[3 * slot + 2] = c;
    }
    });
    phase("leaf slots");
    auto track = [&](void *p) { if (p) s->deviceAllocs.push_back(p); };
#define UP(vec, field, T)                                        \
    {                                                            \
        T *ptr = nullptr;                                        \
        e = uploadVec(vec, &ptr);                                \
        if (e != cudaSuccess) {                                  \
            IileSceneFree(s);                                    \
            return cudaFail(e, "scene upload (" #field ")");     \
        }                                                        \
        track(ptr);                                              \
        d.field = reinterpret_cast<decltype(d.field)>(ptr);      \
    }
    static_assert(sizeof(IileBVHNode) == 32, "LinearBVHNode must be 32 bytes");
    // The reference's binary LinearBVHNode array reaches the device as the wide nodes built from it below (two
    // binary levels per record); no kernel walks the binary form, so it stays on the host (iile_scene_get_nodes,
    // flat files) instead of taking 32 B per node of HBM and a third of the upload time of a large scene.
    d.nodes = nullptr;
    {
        std::vector<WideNode> wide;
        int depth = 0;
        buildWideNodes(h.nodes, &wide, &depth);
        relabelTopBreadthFirst(&wide, kTopStaged);
        phase("wide nodes");
        // a wide level pops one entry and pushes at most four
        if (3 * depth + 2 > 96) {
            IileSceneFree(s);
            return setError(IILE_E_UNSUPPORTED, "BVH too deep for the traversal stack (the reference's own "
                                                "nodesToVisit[64] would overflow as well)");
        }
        UP(wide, wideNodes, WideNode)
        d.nWide = (uint32_t)wide.size();
    }
    UP(leaf, leafPrims, float4)
    UP(h.prims, prims, IilePrim)
    UP(h.P, P, float)
    UP(h.N, N, float)
    UP(h.UV, UV, float)
    UP(h.spheres, spheres, IileSphere)
    UP(h.materials, materials, IileMaterial)
    UP(h.lights, lights, IileLight)
    UP(h.textures, textures, IileTexture)
    UP(h.images, images, IileImage)
    UP(h.texels, texels, float)
    UP(h.materialTextures, materialTextures, IileMaterialTextures)
    if (!h.textures.empty()) {
        std::vector<float> lut(128);  // mipmap.h:197-203
        for (int i = 0; i < 128; ++i) {
            const float alpha = 2, r2 = float(i) / float(128 - 1);
            lut[i] = std::exp(-alpha * r2) - std::exp(-alpha);
        }
        UP(lut, ewaWeights, float)
    }
#undef UP
    phase("copies");
    {
        uint32_t *ctr = nullptr;
        e = cudaMalloc((void **)&ctr, kCounterPairs * 2 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(ctr, 0, kCounterPairs * 2 * sizeof(uint32_t));
        if (e != cudaSuccess) {
            IileSceneFree(s);
            return cudaFail(e, "scene upload (work counters)");
        }
        track(ctr);
        d.castCounters = ctr;
    }
    d.nNodes = (uint32_t)h.nodes.size();
    d.nPrims = (uint32_t)h.prims.size();
    d.nSpheres = (uint32_t)h.spheres.size();
    d.nMaterials = (uint32_t)h.materials.size();
    d.nLights = (uint32_t)h.lights.size();
    d.nInfinite = 0;
    for (const IileLight &l : h.lights) d.nInfinite += l.type == IILE_LIGHT_INFINITE ? 1u : 0u;
    // what the shade launches of this scene need (render_api.cu): the shading classes that occur, and
    // whether every light is a point light or a diffuse area light on a triangle
    s->classesPresent = 1;  // kClassMiss
    s->hasInterfaces = false;
    for (const IilePrim &p : h.prims) s->hasInterfaces |= p.material < 0;
    for (const IilePrim &p : h.prims)
        s->classesPresent |= 1 << ((p.type == IILE_PRIM_TRIANGLE && p.material >= 0 && p.material < (int)h.materials.size() &&
                                    !h.materials[p.material].texSet)
                                       ? std::min(h.materials[p.material].type + 1, 5) : 5);
    s->simpleLights = true;
    for (const IileLight &l : h.lights)
        if (!(l.type == IILE_LIGHT_POINT ||
              (l.type == IILE_LIGHT_DIFFUSE_AREA && l.prim >= 0 && h.prims[l.prim].type == IILE_PRIM_TRIANGLE)))
            s->simpleLights = false;
    for (int k = 0; k < 3; ++k) {
        d.worldMin[k] = h.worldBound.pMin[k];
        d.worldMax[k] = h.worldBound.pMax[k];
    }
    s->uploaded = true;
    s->lightTablesReady = false;
    return IILE_OK;
}

}  // namespace iile

#define IILE_TRY try {
#define IILE_CATCH(code)                                                       \
    }                                                                          \
    catch (const std::exception &ex) { return setError(code, ex.what()); }     \
    catch (...) { return setError(code, "unknown exception"); }

static int requireDevice(IileScene *scene, const char *who) {
    if (!scene) return setError(IILE_E_INVALID, std::string(who) + ": null scene");
    if (!scene->uploaded) {
        int rc = uploadScene(scene);
        if (rc != IILE_OK) return rc;
    }
    return IILE_OK;
}

// Host-buffer variants.  The batch is cut into chunks that flow through three stream-ordered
// lanes (H2D of chunk k+1, traversal of chunk k and D2H of chunk k-1 overlap on the two copy
// engines and the SMs), each lane with its own device staging buffers and its own pair of
// work counters.  The staging area lives as long as the scene.
namespace {
constexpr int kLanes = 3;
constexpr size_t kLaneRays = (size_t)1 << 20;  // 32 MiB of rays, 16 MiB of hit records per lane

struct CastStaging {
    cudaStream_t stream[kLanes] = {};
    IileRay *rays[kLanes] = {};
    void *out[kLanes] = {};
    bool ready = false;
};

cudaError_t ensureStaging(IileScene *s, CastStaging **out) {
    CastStaging *st = static_cast<CastStaging *>(s->castStaging);
    if (!st) {
        st = new CastStaging();
        s->castStaging = st;
    }
    *out = st;
    if (st->ready) return cudaSuccess;
    for (int i = 0; i < kLanes; ++i) {  // a retry after a failed attempt keeps what it already has
        cudaError_t e = cudaSuccess;
        if (!st->stream[i]) e = cudaStreamCreateWithFlags(&st->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess && !st->rays[i]) e = cudaMalloc((void **)&st->rays[i], kLaneRays * sizeof(IileRay));
        if (e == cudaSuccess && !st->out[i]) e = cudaMalloc(&st->out[i], kLaneRays * sizeof(IileHit));
        if (e != cudaSuccess) return e;
    }
    st->ready = true;
    return cudaSuccess;
}
}  // namespace

namespace iile {
void freeCastStaging(IileScene *s) {
    CastStaging *st = static_cast<CastStaging *>(s->castStaging);
    if (!st) return;
    for (int i = 0; i < kLanes; ++i) {
        if (st->rays[i]) cudaFree(st->rays[i]);
        if (st->out[i]) cudaFree(st->out[i]);
        if (st->stream[i]) cudaStreamDestroy(st->stream[i]);
    }
    delete st;
    s->castStaging = nullptr;
}
}  // namespace iile

template <typename OutT, typename Launch>
static int castHost(IileScene *scene, const IileRay *rays, size_t n, OutT *out, Launch launch,
                    const char *who) {
    int rc = requireDevice(scene, who);
    if (rc != IILE_OK) return rc;
    if (n == 0) return IILE_OK;
    if (!rays || !out) return setError(IILE_E_INVALID, std::string(who) + ": null buffer");
    CastStaging *st = nullptr;
    cudaError_t e = ensureStaging(scene, &st);
    if (e != cudaSuccess) return cudaFail(e, "cast staging allocation");
    int lane = 0;
    for (size_t off = 0; off < n && e == cudaSuccess; off += kLaneRays, lane = (lane + 1) % kLanes) {
        const size_t m = std::min(kLaneRays, n - off);
        cudaStream_t sm = st->stream[lane];
        DeviceScene dev;  // this lane's stream has its own {work, done} counters
        if ((rc = deviceViewFor(scene, sm, &dev)) != IILE_OK) break;
        e = cudaMemcpyAsync(st->rays[lane], rays + off, m * sizeof(IileRay), cudaMemcpyHostToDevice, sm);
        if (e == cudaSuccess) e = launch(dev, st->rays[lane], m, static_cast<OutT *>(st->out[lane]), sm);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(out + off, st->out[lane], m * sizeof(OutT), cudaMemcpyDeviceToHost, sm);
    }
    for (int i = 0; i < kLanes; ++i) {
        cudaError_t e2 = cudaStreamSynchronize(st->stream[i]);
        if (e == cudaSuccess) e = e2;
    }
    if (rc != IILE_OK) return rc;
    if (e != cudaSuccess) return cudaFail(e, who);
    return IILE_OK;
}

extern "C" {

const char *iile_last_error(void) { return g_lastError.c_str(); }
int iile_abi_version(void) { return IILE_ABI_VERSION; }

int iile_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int iile_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cudaFail(e, "cudaSetDevice");
    return IILE_OK;
}

static int finishCreate(IileScene *s, int upload, IileScene **out) {
    if (upload) {
        int rc = uploadScene(s);
        if (rc != IILE_OK) {
            delete s;
            return rc;
        }
    }
    *out = s;
    return IILE_OK;
}

int iile_scene_load_pbrt(const char *path, int upload, IileScene **out) {
    if (!path || !out) return setError(IILE_E_INVALID, "iile_scene_load_pbrt: null argument");
    *out = nullptr;
    IileScene *s = new IileScene();
    const auto t0 = std::chrono::steady_clock::now();
    try {
        loadPbrtFile(path, &s->host);
    } catch (const std::exception &ex) {
        delete s;
        return setError(IILE_E_PARSE, ex.what());
    }
    const auto t1 = std::chrono::steady_clock::now();
    const double bvhSeconds = s->host.buildSeconds;  // finishCreate deletes s when the upload fails
    const int rc = finishCreate(s, upload, out);
    if (std::getenv("IILE_BUILD_TRACE"))
        std::fprintf(stderr, "load: parse + flatten + BVH %.3f s (BVH %.3f s), device layout + upload %.3f s\n",
                     std::chrono::duration<double>(t1 - t0).count(), bvhSeconds,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    return rc;
}

int iile_scene_create(const IileSceneDesc *desc, int upload, IileScene **out) {
    if (!desc || !out) return setError(IILE_E_INVALID, "iile_scene_create: null argument");
    *out = nullptr;
    IileScene *s = new IileScene();
    try {
        fromDesc(*desc, &s->host);
    } catch (const std::exception &ex) {
        delete s;
        return setError(IILE_E_INVALID, ex.what());
    }
    return finishCreate(s, upload, out);
}

int iile_scene_load_flat(const char *path, int upload, IileScene **out) {
    if (!path || !out) return setError(IILE_E_INVALID, "iile_scene_load_flat: null argument");
    *out = nullptr;
    IileScene *s = new IileScene();
    try {
        loadFlat(path, &s->host);
    } catch (const std::exception &ex) {
        delete s;
        return setError(IILE_E_IO, ex.what());
    }
    return finishCreate(s, upload, out);
}

int iile_scene_save_flat(const IileScene *scene, const char *path) {
    if (!scene || !path) return setError(IILE_E_INVALID, "iile_scene_save_flat: null argument");
    IILE_TRY
    saveFlat(scene->host, path);
    return IILE_OK;
    IILE_CATCH(IILE_E_IO)
}

void iile_scene_destroy(IileScene *scene) {
    if (!scene) return;
    IileSceneFree(scene);
    delete scene;
}

int iile_scene_info(const IileScene *scene, IileSceneInfo *out) {
    if (!scene || !out) return setError(IILE_E_INVALID, "iile_scene_info: null argument");
    const HostScene &h = scene->host;
    std::memset(out, 0, sizeof(*out));
    out->nVerts = h.nVerts();
    out->nPrims = (uint32_t)h.prims.size();
    out->nSpheres = (uint32_t)h.spheres.size();
    out->nTriangles = out->nPrims - out->nSpheres;
    out->nMaterials = (uint32_t)h.materials.size();
    out->nLights = (uint32_t)h.lights.size();
    out->nNodes = (uint32_t)h.nodes.size();
    for (int k = 0; k < 3; ++k) {
        out->worldMin[k] = h.worldBound.pMin[k];
        out->worldMax[k] = h.worldBound.pMax[k];
    }
    out->setup = h.setup;
    FilmGeometry fg = filmGeometry(h.setup);
    std::memcpy(out->sampleBounds, fg.sampleBounds, sizeof(fg.sampleBounds));
    std::memcpy(out->croppedBounds, fg.cropped, sizeof(fg.cropped));
    out->buildSeconds = h.buildSeconds;
    return IILE_OK;
}

int iile_scene_get_nodes(const IileScene *scene, IileBVHNode *nodes, uint32_t cap) {
    if (!scene || !nodes) return setError(IILE_E_INVALID, "iile_scene_get_nodes: null argument");
    if (cap < scene->host.nodes.size()) return setError(IILE_E_INVALID, "iile_scene_get_nodes: buffer too small");
    std::memcpy(nodes, scene->host.nodes.data(), scene->host.nodes.size() * sizeof(IileBVHNode));
    return IILE_OK;
}

int iile_scene_get_prim_order(const IileScene *scene, uint32_t *order, uint32_t cap) {
    if (!scene || !order) return setError(IILE_E_INVALID, "iile_scene_get_prim_order: null argument");
    if (cap < scene->host.primOrder.size()) return setError(IILE_E_INVALID, "iile_scene_get_prim_order: buffer too small");
    std::memcpy(order, scene->host.primOrder.data(), scene->host.primOrder.size() * 4);
    return IILE_OK;
}

int iile_scene_set_setup(IileScene *scene, const IileRenderSetup *setup) {
    if (!scene || !setup) return setError(IILE_E_INVALID, "iile_scene_set_setup: null argument");
    try {
        IileRenderSetup check = *setup;  // accelerator fields are not taken over: validate them as they stay
        check.maxNodePrims = scene->host.setup.maxNodePrims;
        check.splitMethod = scene->host.setup.splitMethod;
        validateSetup(check);
    } catch (const std::exception &ex) {
        return setError(IILE_E_INVALID, std::string("iile_scene_set_setup: ") + ex.what());
    }
    bool rebuildLights = setup->lightSampleStrategy != scene->host.setup.lightSampleStrategy;
    // accelerator parameters are fixed at scene creation
    int mnp = scene->host.setup.maxNodePrims, sm = scene->host.setup.splitMethod;
    scene->host.setup = *setup;
    scene->host.setup.maxNodePrims = mnp;
    scene->host.setup.splitMethod = sm;
    if (rebuildLights) scene->lightTablesReady = false;
    return IILE_OK;
}

// ---- ray casting ------------------------------------------------------------
int iile_intersect_closest_dev(IileScene *scene, const IileRay *rays, size_t n, IileHit *hits,
                               void *stream) {
    int rc = requireDevice(scene, "iile_intersect_closest_dev");
    if (rc != IILE_OK) return rc;
    if (n && (!rays || !hits)) return setError(IILE_E_INVALID, "iile_intersect_closest_dev: null buffer");
    DeviceScene dev;
    if ((rc = deviceViewFor(scene, (cudaStream_t)stream, &dev)) != IILE_OK) return rc;
    cudaError_t e = launchClosestHit(dev, rays, n, hits, (cudaStream_t)stream);
    if (e != cudaSuccess) return cudaFail(e, "closest-hit launch");
    return IILE_OK;
}

int iile_intersect_any_dev(IileScene *scene, const IileRay *rays, size_t n, uint8_t *occluded,
                           void *stream) {
    int rc = requireDevice(scene, "iile_intersect_any_dev");
    if (rc != IILE_OK) return rc;
    if (n && (!rays || !occluded)) return setError(IILE_E_INVALID, "iile_intersect_any_dev: null buffer");
    DeviceScene dev;
    if ((rc = deviceViewFor(scene, (cudaStream_t)stream, &dev)) != IILE_OK) return rc;
    cudaError_t e = launchAnyHit(dev, rays, n, occluded, (cudaStream_t)stream);
    if (e != cudaSuccess) return cudaFail(e, "any-hit launch");
    return IILE_OK;
}

int iile_intersect_closest(IileScene *scene, const IileRay *rays, size_t n, IileHit *hits) {
    return castHost(scene, rays, n, hits, launchClosestHit, "iile_intersect_closest");
}

int iile_intersect_any(IileScene *scene, const IileRay *rays, size_t n, uint8_t *occluded) {
    return castHost(scene, rays, n, occluded, launchAnyHit, "iile_intersect_any");
}

int iile_generate_rays_dev(const float bmin[3], const float bmax[3], uint64_t firstRay, size_t n,
                           IileRay *rays, void *stream) {
    if (!bmin || !bmax || (n && !rays)) return setError(IILE_E_INVALID, "iile_generate_rays_dev: null argument");
    if (iile_device_count() == 0) return setError(IILE_E_CUDA, "no CUDA device available");
    cudaError_t e = launchGenerateRays(bmin, bmax, firstRay, n, rays, (cudaStream_t)stream);
    if (e != cudaSuccess) return cudaFail(e, "ray generator launch");
    return IILE_OK;
}

}  // extern "C"
