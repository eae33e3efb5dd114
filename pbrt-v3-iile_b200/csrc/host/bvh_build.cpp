// Host-side construction of the LinearBVHNode array the traversal kernels walk.
//
// Same tree as the reference's BVHAccel (src/accelerators/bvh.cpp:183-402):
// centroid-bounds maximum-extent axis, 12-bucket SAH with cost
// 1 + (n0*A0 + n1*A1)/A, leaf when n <= maxPrimsInNode and leafCost <= minCost,
// nth_element median for n <= 2, degenerate-centroid leaves, `middle` /
// `equal` split methods; flattened depth-first with the first child at
// index+1 (bvh.cpp:640-658).  The build is iterative over an explicit work
// stack rather than recursive, and runs on all host cores (see Builder): the
// result is byte-identical for any thread count.
//
// Two details are pinned to the reference *as compiled by g++*, because they
// decide the primitive order inside the array and hence tie-breaking between
// equal-t hits: (1) InitInterior's two recursiveBuild() calls are function
// arguments, which g++ evaluates right to left, so the SECOND child's subtree
// claims its slots in orderedPrims first; (2) std::partition / std::nth_element
// are libstdc++'s.  oracle/ref_harness dumps the reference's own arrays and
// tests/ compares them bit for bit.
#include <cmath>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>

#include "host_scene.h"

namespace iile {

Bounds3 primWorldBound(const HostScene &s, const IilePrim &p) {
    if (p.type == IILE_PRIM_TRIANGLE) {  // triangle.cpp:176-186
        return unite(Bounds3(s.vertex(p.v[0]), s.vertex(p.v[1])), s.vertex(p.v[2]));
    }
    const IileSphere &sp = s.spheres[p.v[0]];  // sphere.cpp:44-47 + Shape::WorldBound
    Xform o2w;
    std::memcpy(o2w.m.m, sp.objectToWorld, sizeof(float) * 16);
    std::memcpy(o2w.inv.m, sp.worldToObject, sizeof(float) * 16);
    if (sp.flags & 4u)  // Disk::ObjectBound, disk.cpp:43-46 (height in zMin)
        return o2w.bounds(Bounds3(V3(-sp.radius, -sp.radius, sp.zMin), V3(sp.radius, sp.radius, sp.zMin)));
    return o2w.bounds(Bounds3(V3(-sp.radius, -sp.radius, sp.zMin), V3(sp.radius, sp.radius, sp.zMax)));
}

float primArea(const HostScene &s, const IilePrim &p) {
    if (p.type == IILE_PRIM_TRIANGLE) {  // triangle.cpp:546-552
        V3 p0 = s.vertex(p.v[0]), p1 = s.vertex(p.v[1]), p2 = s.vertex(p.v[2]);
        return 0.5f * length(cross(p1 - p0, p2 - p0));
    }
    const IileSphere &sp = s.spheres[p.v[0]];  // sphere.cpp:217; disk.cpp:131-133 (the 0.5 is a double there)
    if (sp.flags & 4u) return (float)(sp.phiMax * 0.5 * (sp.radius * sp.radius - sp.zMax * sp.zMax));
    if (sp.flags & 8u) return (sp.zMax - sp.zMin) * sp.radius * sp.phiMax;  // cylinder.cpp:202
    return sp.phiMax * sp.radius * (sp.zMax - sp.zMin);
}

namespace {

struct PrimInfo {
    size_t number;
    Bounds3 bounds;
    V3 centroid;
};

struct Bucket {
    int count = 0;
    Bounds3 bounds;
};

constexpr int kBuckets = 12;

// A node under construction; children are linked by index for the final flatten.  Trivially
// constructible on purpose: the node pool is allocated untouched (2n slots) and filled as nodes
// are claimed.
struct BuildNode {
    Bounds3 bounds;
    int child[2];
    int axis, firstPrim, nPrims;
    int count;  // nodes in this subtree (set when the subtree is complete): slots of the flat array
};

// Run fn(threadIndex, begin, end) over [start,end) cut into `parts` contiguous pieces.
template <typename F>
void parallelChunks(int start, int end, int parts, F fn) {
    if (parts <= 1) {
        fn(0, start, end);
        return;
    }
    std::vector<std::thread> th;
    const int64_t n = end - start;
    for (int t = 0; t < parts; ++t) {
        const int b = start + (int)(n * t / parts), e = start + (int)(n * (t + 1) / parts);
        th.emplace_back([=] { fn(t, b, e); });
    }
    for (auto &x : th) x.join();
}

// The build is parallel in two ways that cannot change the result (the reference's tree is a pure
// function of the primitive bounds and of libstdc++'s partition / nth_element on each range):
//   * the reductions of a large range — node bounds, centroid bounds, SAH buckets — are float
//     min/max and integer counts, exact in any order, and are cut across threads; the partition of
//     the range itself stays one std::partition call;
//   * ranges are independent once split, so the top of the tree is expanded on the calling thread
//     down to subtrees of at most `grain` primitives, which worker threads then build concurrently.
// A leaf's slots in the ordered primitive array follow from its range alone: with the second child
// built first (header comment) the leaves are created by decreasing range, so the leaf of
// [start,end) begins at n - end.
struct Builder {
    std::vector<PrimInfo> info;
    std::unique_ptr<BuildNode[]> tree;
    std::atomic<int> nextNode{0};
    std::vector<uint32_t> ordered;
    int maxPrims;
    int method;
    int nTotal = 0;
    int nThreads = 1;
    double topSeconds = 0, subSeconds = 0;  // IILE_BUILD_TRACE
    static constexpr int kParallelRange = 1 << 17;  // ranges this large cut their reductions across threads

    int newNode() {
        const int i = nextNode.fetch_add(1, std::memory_order_relaxed);
        BuildNode &nd = tree[i];
        nd.child[0] = nd.child[1] = -1;
        nd.axis = nd.firstPrim = nd.nPrims = 0;
        return i;
    }
    int partsFor(int n) const { return n >= kParallelRange ? nThreads : 1; }

    int leaf(int node, int start, int end, const Bounds3 &b) {
        tree[node].firstPrim = nTotal - end;
        tree[node].nPrims = end - start;
        tree[node].bounds = b;
        tree[node].count = 1;
        for (int i = start; i < end; ++i) ordered[(size_t)(nTotal - end) + (i - start)] = (uint32_t)info[i].number;
        return node;
    }

    // bounds of the primitives and of their centroids over [start,end), one pass
    void rangeBounds(int start, int end, Bounds3 *bounds, Bounds3 *centroidBounds) const {
        const int parts = partsFor(end - start);
        std::vector<Bounds3> part(2 * (size_t)parts);
        parallelChunks(start, end, parts, [&](int t, int b, int e) {
            Bounds3 acc, cen;
            for (int i = b; i < e; ++i) {
                acc = unite(acc, info[i].bounds);
                cen = unite(cen, info[i].centroid);
            }
            part[2 * t] = acc;
            part[2 * t + 1] = cen;
        });
        Bounds3 ob, oc;
        for (int t = 0; t < parts; ++t) {
            ob = unite(ob, part[2 * t]);
            oc = unite(oc, part[2 * t + 1]);
        }
        *bounds = ob;
        *centroidBounds = oc;
    }

    // std::partition of info[start,end) — libstdc++'s bidirectional algorithm (stl_algo.h
    // __partition): the i-th misplaced element from the left (predicate false inside the first
    // `mid` slots, ascending) is swapped with the i-th misplaced element from the right (predicate
    // true behind them, descending), everything else stays where it is.  The same permutation is
    // produced here by counting and pairing in parallel; small ranges call std::partition itself.
    template <typename Pred>
    int partitionRange(int start, int end, Pred pred, int knownTrues = -1) {
        const int parts = partsFor(end - start);
        if (parts <= 1) {
            PrimInfo *m = std::partition(&info[start], &info[end - 1] + 1, pred);
            return (int)(m - &info[0]);
        }
        int total = knownTrues;
        if (total < 0) {
            std::vector<int> trues(parts);
            parallelChunks(start, end, parts, [&](int t, int b, int e) {
                int c = 0;
                for (int i = b; i < e; ++i) c += pred(info[i]) ? 1 : 0;
                trues[t] = c;
            });
            total = 0;
            for (int c : trues) total += c;
        }
        const int mid = start + total;
        if (mid == start || mid == end) return mid;
        // misplaced on the left: pred false in [start,mid); on the right: pred true in [mid,end)
        std::vector<std::vector<int>> leftOf(parts), rightOf(parts);
        parallelChunks(start, end, parts, [&](int t, int b, int e) {
            std::vector<int> &l = leftOf[t], &r = rightOf[t];
            for (int i = b; i < std::min(e, mid); ++i)
                if (!pred(info[i])) l.push_back(i);
            for (int i = std::max(b, mid); i < e; ++i)
                if (pred(info[i])) r.push_back(i);
        });
        std::vector<int> left, right;  // both ascending; `right` is consumed from the back
        for (int t = 0; t < parts; ++t) {
            left.insert(left.end(), leftOf[t].begin(), leftOf[t].end());
            right.insert(right.end(), rightOf[t].begin(), rightOf[t].end());
        }
        const int nSwap = (int)left.size();  // == right.size(): as many falses before mid as trues behind it
        parallelChunks(0, nSwap, parts, [&](int, int b, int e) {
            for (int k = b; k < e; ++k) std::swap(info[left[k]], info[right[nSwap - 1 - k]]);
        });
        return mid;
    }

    // Decide the split of [start,end).  Returns -1 for "make a leaf".
    int split(int start, int end, const Bounds3 &bounds, const Bounds3 &cb, int *dimOut) {
        int n = end - start;
        int dim = cb.maximumExtent();
        *dimOut = dim;
        if (cb.pMax[dim] == cb.pMin[dim]) return -1;
        int mid = (start + end) / 2;
        auto byCentroid = [dim](const PrimInfo &a, const PrimInfo &b) {
            return a.centroid[dim] < b.centroid[dim];
        };
        if (method == IILE_SPLIT_MIDDLE) {
            float pmid = (cb.pMin[dim] + cb.pMax[dim]) / 2;
            mid = partitionRange(start, end, [dim, pmid](const PrimInfo &pi) { return pi.centroid[dim] < pmid; });
            if (mid != start && mid != end) return mid;
            mid = (start + end) / 2;  // falls through to equal counts
            std::nth_element(&info[start], &info[mid], &info[end - 1] + 1, byCentroid);
            return mid;
        }
        if (method == IILE_SPLIT_EQUAL || n <= 2) {
            std::nth_element(&info[start], &info[mid], &info[end - 1] + 1, byCentroid);
            return mid;
        }
        // SAH
        Bucket buckets[kBuckets];
        auto bucketOf = [&](const PrimInfo &pi) {
            int b = (int)(kBuckets * cb.offset(pi.centroid)[dim]);
            if (b == kBuckets) b = kBuckets - 1;
            return b;
        };
        {
            const int parts = partsFor(n);
            std::vector<std::array<Bucket, kBuckets>> part(parts);
            parallelChunks(start, end, parts, [&](int t, int b0, int e0) {
                std::array<Bucket, kBuckets> &bk = part[t];
                for (int i = b0; i < e0; ++i) {
                    int b = bucketOf(info[i]);
                    bk[b].count++;
                    bk[b].bounds = unite(bk[b].bounds, info[i].bounds);
                }
            });
            for (const auto &bk : part)
                for (int b = 0; b < kBuckets; ++b) {
                    buckets[b].count += bk[b].count;
                    buckets[b].bounds = unite(buckets[b].bounds, bk[b].bounds);
                }
        }
        float cost[kBuckets - 1];
        for (int i = 0; i < kBuckets - 1; ++i) {
            Bounds3 b0, b1;
            int c0 = 0, c1 = 0;
            for (int j = 0; j <= i; ++j) {
                b0 = unite(b0, buckets[j].bounds);
                c0 += buckets[j].count;
            }
            for (int j = i + 1; j < kBuckets; ++j) {
                b1 = unite(b1, buckets[j].bounds);
                c1 += buckets[j].count;
            }
            cost[i] = 1 + (c0 * b0.surfaceArea() + c1 * b1.surfaceArea()) / bounds.surfaceArea();
        }
        float minCost = cost[0];
        int minBucket = 0;
        for (int i = 1; i < kBuckets - 1; ++i)
            if (cost[i] < minCost) {
                minCost = cost[i];
                minBucket = i;
            }
        float leafCost = (float)n;
        if (n > maxPrims || minCost < leafCost) {
            int below = 0;
            for (int j = 0; j <= minBucket; ++j) below += buckets[j].count;
            return partitionRange(start, end, [&](const PrimInfo &pi) { return bucketOf(pi) <= minBucket; }, below);
        }
        return -1;
    }

    // Depth-first build of the subtree rooted at `root` with an explicit stack.  Each frame is
    // visited twice: first to split and schedule its children, then to merge their bounds.
    // Ranges of at most `deferBelow` primitives (other than the root's own) are not built but
    // appended to `deferred`; their ancestors' bounds are then left for mergeTop().
    struct Frame {
        int node, start, end, stage;
    };
    void build(Frame root, int deferBelow, std::vector<Frame> *deferred, std::vector<int> *topNodes) {
        std::vector<Frame> stack;
        stack.push_back(root);
        while (!stack.empty()) {
            Frame fr = stack.back();
            stack.pop_back();
            if (fr.stage == 1) {
                BuildNode &nd = tree[fr.node];
                nd.bounds = unite(tree[nd.child[0]].bounds, tree[nd.child[1]].bounds);
                nd.count = 1 + tree[nd.child[0]].count + tree[nd.child[1]].count;
                continue;
            }
            int n = fr.end - fr.start;
            if (deferred && n <= deferBelow) {
                deferred->push_back(fr);
                continue;
            }
            Bounds3 bounds, cb;
            rangeBounds(fr.start, fr.end, &bounds, &cb);
            if (n == 1) {
                leaf(fr.node, fr.start, fr.end, bounds);
                continue;
            }
            int dim = 0;
            int mid = split(fr.start, fr.end, bounds, cb, &dim);
            if (mid < 0) {
                leaf(fr.node, fr.start, fr.end, bounds);
                continue;
            }
            if (mid == fr.start || mid == fr.end)
                throw std::runtime_error("BVH build: empty partition (the reference would fail its CHECK_NE here)");
            const int c0 = newNode(), c1 = newNode();
            tree[fr.node].axis = dim;
            tree[fr.node].child[0] = c0;
            tree[fr.node].child[1] = c1;
            if (topNodes) topNodes->push_back(fr.node);  // bounds merged after the deferred subtrees exist
            else stack.push_back({fr.node, fr.start, fr.end, 1});
            stack.push_back({c0, fr.start, mid, 0});
            stack.push_back({c1, mid, fr.end, 0});
        }
    }

    // ---- splitmethod "hlbvh" (bvh.cpp:404-638) ------------------------------------------------
    // Morton codes of the centroids (10 bits per axis), LSD radix sort, one LBVH treelet per
    // distinct top-12-bit prefix, SAH over the treelet roots.  The reference hands leaves their
    // slots in orderedPrims through an atomic counter, so with several threads its primitive
    // order depends on scheduling; single-threaded it is the Morton-sorted order, which is what
    // this builder always produces (any of the reference's orders is a valid tree; this one is
    // reproducible).  Treelets are emitted concurrently.
    struct MortonPrim {
        int prim;
        uint32_t code;
    };
    static uint32_t leftShift3(uint32_t x) {
        if (x == (1u << 10)) --x;
        x = (x | (x << 16)) & 0x30000ffu;
        x = (x | (x << 8)) & 0x300f00fu;
        x = (x | (x << 4)) & 0x30c30c3u;
        x = (x | (x << 2)) & 0x9249249u;
        return x;
    }
    int emitLBVH(const MortonPrim *mp, int base, int n, int bit) {
        if (bit == -1 || n < maxPrims) {
            const int node = newNode();
            Bounds3 b;
            for (int i = 0; i < n; ++i) b = unite(b, info[mp[i].prim].bounds);
            tree[node].firstPrim = base;
            tree[node].nPrims = n;
            tree[node].bounds = b;
            tree[node].count = 1;
            return node;
        }
        const uint32_t mask = 1u << bit;
        if ((mp[0].code & mask) == (mp[n - 1].code & mask)) return emitLBVH(mp, base, n, bit - 1);
        int lo = 0, hi = n - 1;
        while (lo + 1 != hi) {
            const int mid = (lo + hi) / 2;
            if ((mp[lo].code & mask) == (mp[mid].code & mask)) lo = mid;
            else hi = mid;
        }
        const int split = hi;
        const int node = newNode();
        const int c0 = emitLBVH(mp, base, split, bit - 1);
        const int c1 = emitLBVH(mp + split, base + split, n - split, bit - 1);
        BuildNode &nd = tree[node];
        nd.axis = bit % 3;
        nd.child[0] = c0;
        nd.child[1] = c1;
        nd.bounds = unite(tree[c0].bounds, tree[c1].bounds);
        nd.count = 1 + tree[c0].count + tree[c1].count;
        return node;
    }
    int upperSAH(std::vector<int> &roots, int start, int end) {
        if (end - start == 1) return roots[start];
        Bounds3 bounds, cb;
        for (int i = start; i < end; ++i) {
            const Bounds3 &b = tree[roots[i]].bounds;
            bounds = unite(bounds, b);
            cb = unite(cb, (b.pMin + b.pMax) * 0.5f);
        }
        const int dim = cb.maximumExtent();
        if (cb.pMax[dim] == cb.pMin[dim])
            throw std::runtime_error("HLBVH: treelet centroids coincide (the reference fails its CHECK_NE here)");
        auto bucketOf = [&](int root) {
            const Bounds3 &b = tree[root].bounds;
            const float centroid = (b.pMin[dim] + b.pMax[dim]) * 0.5f;
            int k = (int)(kBuckets * ((centroid - cb.pMin[dim]) / (cb.pMax[dim] - cb.pMin[dim])));
            if (k == kBuckets) k = kBuckets - 1;
            return k;
        };
        Bucket buckets[kBuckets];
        for (int i = start; i < end; ++i) {
            const int k = bucketOf(roots[i]);
            buckets[k].count++;
            buckets[k].bounds = unite(buckets[k].bounds, tree[roots[i]].bounds);
        }
        float cost[kBuckets - 1];
        for (int i = 0; i < kBuckets - 1; ++i) {
            Bounds3 b0, b1;
            int c0 = 0, c1 = 0;
            for (int j = 0; j <= i; ++j) {
                b0 = unite(b0, buckets[j].bounds);
                c0 += buckets[j].count;
            }
            for (int j = i + 1; j < kBuckets; ++j) {
                b1 = unite(b1, buckets[j].bounds);
                c1 += buckets[j].count;
            }
            cost[i] = .125f + (c0 * b0.surfaceArea() + c1 * b1.surfaceArea()) / bounds.surfaceArea();
        }
        float minCost = cost[0];
        int minBucket = 0;
        for (int i = 1; i < kBuckets - 1; ++i)
            if (cost[i] < minCost) {
                minCost = cost[i];
                minBucket = i;
            }
        int *m = std::partition(&roots[start], &roots[end - 1] + 1, [&](int r) { return bucketOf(r) <= minBucket; });
        const int mid = (int)(m - &roots[0]);
        if (mid == start || mid == end)
            throw std::runtime_error("HLBVH: empty SAH partition over the treelets (the reference fails its CHECK here)");
        const int node = newNode();
        const int c0 = upperSAH(roots, start, mid), c1 = upperSAH(roots, mid, end);
        BuildNode &nd = tree[node];
        nd.axis = dim;
        nd.child[0] = c0;
        nd.child[1] = c1;
        nd.bounds = unite(tree[c0].bounds, tree[c1].bounds);
        nd.count = 1 + tree[c0].count + tree[c1].count;
        return node;
    }
    int rootNode = 0;
    int buildHLBVH(int nPrims) {
        nTotal = nPrims;
        tree.reset(new BuildNode[2 * (size_t)nPrims + 8192]);
        ordered.resize(nPrims);
        Bounds3 all, cb;
        rangeBounds(0, nPrims, &all, &cb);
        std::vector<MortonPrim> mp(nPrims), tmp(nPrims);
        parallelChunks(0, nPrims, partsFor(nPrims), [&](int, int b, int e) {
            for (int i = b; i < e; ++i) {
                const V3 o = cb.offset(info[i].centroid) * 1024.f;
                mp[i].prim = (int)info[i].number;
                mp[i].code = (leftShift3((uint32_t)o.z) << 2) | (leftShift3((uint32_t)o.y) << 1) | leftShift3((uint32_t)o.x);
            }
        });
        for (int pass = 0; pass < 5; ++pass) {  // RadixSort, bvh.cpp:139-180: 6 bits per pass, stable
            const int lowBit = pass * 6;
            std::vector<MortonPrim> &in = (pass & 1) ? tmp : mp, &out = (pass & 1) ? mp : tmp;
            // per-chunk histograms; a chunk's elements of bucket b land after those of every earlier
            // chunk: the stable order of the sequential pass
            const int parts = partsFor(nPrims);
            std::vector<std::array<int, 64>> count(parts);
            parallelChunks(0, nPrims, parts, [&](int t, int b0, int e0) {
                std::array<int, 64> c{};
                for (int i = b0; i < e0; ++i) ++c[(in[i].code >> lowBit) & 63];
                count[t] = c;
            });
            int acc = 0;
            for (int bkt = 0; bkt < 64; ++bkt)
                for (int t = 0; t < parts; ++t) {
                    const int c = count[t][bkt];
                    count[t][bkt] = acc;
                    acc += c;
                }
            parallelChunks(0, nPrims, parts, [&](int t, int b0, int e0) {
                std::array<int, 64> at = count[t];
                for (int i = b0; i < e0; ++i) out[at[(in[i].code >> lowBit) & 63]++] = in[i];
            });
        }
        mp.swap(tmp);  // five passes: the sorted data ended in tmp
        struct Treelet { int start, n, root; };
        std::vector<Treelet> treelets;
        for (int start = 0, end = 1; end <= nPrims; ++end)
            if (end == nPrims || ((mp[start].code & 0x3ffc0000u) != (mp[end].code & 0x3ffc0000u))) {
                treelets.push_back({start, end - start, -1});
                start = end;
            }
        {
            std::atomic<size_t> next{0};
            std::vector<std::thread> th;
            const int workers = nPrims >= kParallelRange ? nThreads : 1;
            auto work = [&] {
                for (size_t k; (k = next.fetch_add(1)) < treelets.size();)
                    treelets[k].root = emitLBVH(&mp[treelets[k].start], treelets[k].start, treelets[k].n, 29 - 12);
            };
            for (int t = 1; t < workers; ++t) th.emplace_back(work);
            work();
            for (auto &x : th) x.join();
        }
        for (int i = 0; i < nPrims; ++i) ordered[i] = (uint32_t)mp[i].prim;
        std::vector<int> roots;
        for (const Treelet &t : treelets) roots.push_back(t.root);
        rootNode = upperSAH(roots, 0, (int)roots.size());
        return nextNode.load();
    }

    int buildAll(int nPrims) {
        nTotal = nPrims;
        tree.reset(new BuildNode[2 * (size_t)nPrims]);
        ordered.resize(nPrims);
        const int root = newNode();
        if (nThreads <= 1 || nPrims < 4 * kParallelRange) {
            build({root, 0, nPrims, 0}, 0, nullptr, nullptr);
            return nextNode.load();
        }
        // top of the tree on this thread, subtrees of <= grain primitives on the workers
        const int grain = std::max(nPrims / (8 * nThreads), 1 << 14);
        std::vector<Frame> deferred;
        std::vector<int> topNodes;
        auto tt0 = std::chrono::steady_clock::now();
        build({root, 0, nPrims, 0}, grain, &deferred, &topNodes);
        topSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tt0).count();
        tt0 = std::chrono::steady_clock::now();
        std::sort(deferred.begin(), deferred.end(),
                  [](const Frame &a, const Frame &b) { return a.end - a.start > b.end - b.start; });
        std::atomic<size_t> nextTask{0};
        std::vector<std::string> errors(nThreads);
        const int savedThreads = nThreads;
        nThreads = 1;  // inside a worker every range is reduced by the worker alone
        {
            std::vector<std::thread> th;
            for (int t = 0; t < savedThreads; ++t)
                th.emplace_back([&, t] {
                    try {
                        for (size_t k; (k = nextTask.fetch_add(1)) < deferred.size();)
                            build(deferred[k], 0, nullptr, nullptr);
                    } catch (const std::exception &ex) {
                        errors[t] = ex.what();
                    }
                });
            for (auto &x : th) x.join();
        }
        nThreads = savedThreads;
        subSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tt0).count();
        for (const std::string &e : errors)
            if (!e.empty()) throw std::runtime_error(e);
        // children were created after their parents: reverse creation order merges bottom-up
        for (size_t k = topNodes.size(); k-- > 0;) {
            BuildNode &nd = tree[topNodes[k]];
            nd.bounds = unite(tree[nd.child[0]].bounds, tree[nd.child[1]].bounds);
            nd.count = 1 + tree[nd.child[0]].count + tree[nd.child[1]].count;
        }
        return nextNode.load();
    }
};

}  // namespace

// DistantLight::Preprocess (distant.h:55-57): Bounds3::BoundingSphere of the scene bound
// (geometry.h:808-811: centre = (pMin + pMax) / 2 through a reciprocal, radius = Distance(centre, pMax)).
void setLightWorldRadius(HostScene *scene) {
    const Bounds3 &b = scene->worldBound;
    const V3 c = divide(b.pMin + b.pMax, 2.f);
    const bool inside = c.x >= b.pMin.x && c.x <= b.pMax.x && c.y >= b.pMin.y && c.y <= b.pMax.y &&
                        c.z >= b.pMin.z && c.z <= b.pMax.z;
    const float radius = inside ? length(c - b.pMax) : 0.f;
    for (IileLight &l : scene->lights)
        if (l.type == IILE_LIGHT_DISTANT || l.type == IILE_LIGHT_INFINITE) l.worldRadius = radius;
}

DeviceHlbvhFn g_deviceHlbvh = nullptr;

void hlbvhUpperTree(const std::vector<Bounds3> &rootBounds, const std::vector<int> &rootCounts,
                    std::vector<HlbvhUpperNode> *out, int *root) {
    const int nT = (int)rootBounds.size();
    Builder b;
    b.tree.reset(new BuildNode[2 * (size_t)nT + 8]);
    std::vector<int> roots(nT);
    for (int t = 0; t < nT; ++t) {
        const int node = b.newNode();
        b.tree[node].bounds = rootBounds[t];
        b.tree[node].count = rootCounts[t];
        roots[t] = node;
    }
    *root = b.upperSAH(roots, 0, nT);
    const int total = b.nextNode.load();
    out->resize(total);
    for (int i = 0; i < total; ++i) {
        const BuildNode &bn = b.tree[i];
        (*out)[i] = HlbvhUpperNode{bn.bounds, {bn.child[0], bn.child[1]}, bn.axis, bn.count};
    }
}

// Positions the builders (and every kernel after them) can work with: an infinite or NaN coordinate makes the
// centroid bounds infinite and the SAH bucket / Morton index of every primitive undefined (the reference's int
// conversion of a NaN offset is undefined behaviour too; found by tools/fuzz_parser.py as a crash)
static void requireFiniteGeometry(const HostScene &s) {
    for (float v : s.P)
        if (!std::isfinite(v)) throw std::runtime_error("scene: a vertex position is not finite");
    for (const IileSphere &q : s.spheres) {
        bool ok = std::isfinite(q.radius) && std::isfinite(q.zMin) && std::isfinite(q.zMax) && std::isfinite(q.phiMax);
        for (int k = 0; k < 16; ++k) ok = ok && std::isfinite(q.objectToWorld[k]) && std::isfinite(q.worldToObject[k]);
        if (!ok) throw std::runtime_error("scene: a sphere / disk / cylinder has a parameter or transform that is not finite");
    }
}

void buildBVH(HostScene *scene) {
    requireFiniteGeometry(*scene);
    auto t0 = std::chrono::steady_clock::now();
    HostScene &s = *scene;
    size_t n = s.prims.size();
    s.nodes.clear();
    s.primOrder.clear();
    if (n == 0) return;
    if (s.setup.splitMethod == IILE_SPLIT_HLBVH && g_deviceHlbvh && n >= ((size_t)1 << 16)) {
        const char *env = std::getenv("IILE_DEVICE_HLBVH");
        std::string why;
        if ((!env || std::atoi(env) != 0) && g_deviceHlbvh(scene, &why)) {
            s.worldBound = Bounds3(V3(s.nodes[0].bmin[0], s.nodes[0].bmin[1], s.nodes[0].bmin[2]),
                                   V3(s.nodes[0].bmax[0], s.nodes[0].bmax[1], s.nodes[0].bmax[2]));
            setLightWorldRadius(scene);
            s.buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (std::getenv("IILE_BUILD_TRACE")) std::fprintf(stderr, "bvh: HLBVH on the device, %.3f s\n", s.buildSeconds);
            return;
        }
        if (std::getenv("IILE_BUILD_TRACE") && !why.empty())
            std::fprintf(stderr, "bvh: device HLBVH not used (%s); building on the host\n", why.c_str());
        s.nodes.clear();
        s.primOrder.clear();
    }
    Builder b;
    b.maxPrims = std::min(255, s.setup.maxNodePrims);
    b.method = s.setup.splitMethod;
    // host threads: IILE_BUILD_THREADS, else every core (the result does not depend on the count)
    {
        const char *env = std::getenv("IILE_BUILD_THREADS");
        int t = env ? std::atoi(env) : (int)std::thread::hardware_concurrency();
        b.nThreads = std::max(1, std::min(t, 64));
    }
    b.info.resize(n);
    parallelChunks(0, (int)n, n >= (size_t)Builder::kParallelRange ? b.nThreads : 1, [&](int, int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            Bounds3 wb = primWorldBound(s, s.prims[i]);
            b.info[i] = {(size_t)i, wb, 0.5f * wb.pMin + 0.5f * wb.pMax};
        }
    });
    auto tA = std::chrono::steady_clock::now();
    int total = b.method == IILE_SPLIT_HLBVH ? b.buildHLBVH((int)n) : b.buildAll((int)n);
    auto tB = std::chrono::steady_clock::now();
    if (std::getenv("IILE_BUILD_TRACE"))
        std::fprintf(stderr, "bvh: info %.3f s, tree %.3f s (top %.3f, subtrees %.3f)\n",
                     std::chrono::duration<double>(tA - t0).count(), std::chrono::duration<double>(tB - tA).count(),
                     b.topSeconds, b.subSeconds);

    // flatten depth-first, first child adjacent (bvh.cpp:640-658).  A node's slot follows from the
    // subtree sizes — first child at slot+1, second at slot+1+count(first) — so subtrees are
    // flattened concurrently once the top has been walked.
    s.nodes.resize(total);
    struct Item { int build; int slot; };
    auto flatten = [&](Item root, int deferBelow, std::vector<Item> *deferred) {
        std::vector<Item> stack;
        stack.push_back(root);
        while (!stack.empty()) {
            Item it = stack.back();
            stack.pop_back();
            const BuildNode &bn = b.tree[it.build];
            if (deferred && bn.count <= deferBelow) {
                deferred->push_back(it);
                continue;
            }
            IileBVHNode &ln = s.nodes[it.slot];
            std::memset(&ln, 0, sizeof(ln));
            for (int k = 0; k < 3; ++k) {
                ln.bmin[k] = bn.bounds.pMin[k];
                ln.bmax[k] = bn.bounds.pMax[k];
            }
            if (bn.nPrims > 0) {
                if (bn.nPrims >= 65536) throw std::runtime_error("BVH leaf with >= 65536 primitives");
                ln.offset = bn.firstPrim;
                ln.nPrims = (uint16_t)bn.nPrims;
            } else {
                const int second = it.slot + 1 + b.tree[bn.child[0]].count;
                ln.axis = (uint8_t)bn.axis;
                ln.nPrims = 0;
                ln.offset = second;
                stack.push_back({bn.child[1], second});
                stack.push_back({bn.child[0], it.slot + 1});
            }
        }
    };
    if (b.nThreads <= 1 || total < (1 << 18))
        flatten({b.rootNode, 0}, 0, nullptr);
    else {
        std::vector<Item> tasks;
        flatten({b.rootNode, 0}, std::max(total / (8 * b.nThreads), 1 << 12), &tasks);
        std::atomic<size_t> nextTask{0};
        std::vector<std::string> errors(b.nThreads);
        std::vector<std::thread> th;
        for (int t = 0; t < b.nThreads; ++t)
            th.emplace_back([&, t] {
                try {
                    for (size_t k; (k = nextTask.fetch_add(1)) < tasks.size();) flatten(tasks[k], 0, nullptr);
                } catch (const std::exception &ex) {
                    errors[t] = ex.what();
                }
            });
        for (auto &x : th) x.join();
        for (const std::string &e : errors)
            if (!e.empty()) throw std::runtime_error(e);
    }
    s.primOrder = std::move(b.ordered);
    s.worldBound = Bounds3(V3(s.nodes[0].bmin[0], s.nodes[0].bmin[1], s.nodes[0].bmin[2]),
                           V3(s.nodes[0].bmax[0], s.nodes[0].bmax[1], s.nodes[0].bmax[2]));
    setLightWorldRadius(scene);
    s.buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (std::getenv("IILE_BUILD_TRACE"))
        std::fprintf(stderr, "bvh: flatten %.3f s, total %.3f s, %d threads\n",
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - tB).count(), s.buildSeconds, b.nThreads);
}

}  // namespace iile
