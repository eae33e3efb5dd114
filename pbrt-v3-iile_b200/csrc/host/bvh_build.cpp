// Host-side construction of the LinearBVHNode array the traversal kernels walk.
//
// Same tree as the reference's BVHAccel (src/accelerators/bvh.cpp:183-402):
// centroid-bounds maximum-extent axis, 12-bucket SAH with cost
// 1 + (n0*A0 + n1*A1)/A, leaf when n <= maxPrimsInNode and leafCost <= minCost,
// nth_element median for n <= 2, degenerate-centroid leaves, `middle` /
// `equal` split methods; flattened depth-first with the first child at
// index+1 (bvh.cpp:640-658).  The build is iterative over an explicit work
// stack rather than recursive, and writes the flat array directly.
//
// Two details are pinned to the reference *as compiled by g++*, because they
// decide the primitive order inside the array and hence tie-breaking between
// equal-t hits: (1) InitInterior's two recursiveBuild() calls are function
// arguments, which g++ evaluates right to left, so the SECOND child's subtree
// claims its slots in orderedPrims first; (2) std::partition / std::nth_element
// are libstdc++'s.  oracle/ref_harness dumps the reference's own arrays and
// tests/ compares them bit for bit.
#include <chrono>

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
    return o2w.bounds(Bounds3(V3(-sp.radius, -sp.radius, sp.zMin), V3(sp.radius, sp.radius, sp.zMax)));
}

float primArea(const HostScene &s, const IilePrim &p) {
    if (p.type == IILE_PRIM_TRIANGLE) {  // triangle.cpp:546-552
        V3 p0 = s.vertex(p.v[0]), p1 = s.vertex(p.v[1]), p2 = s.vertex(p.v[2]);
        return 0.5f * length(cross(p1 - p0, p2 - p0));
    }
    const IileSphere &sp = s.spheres[p.v[0]];  // sphere.cpp:217
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

// A node under construction.  `slot` is where the finished node's bounds go in
// the temporary tree; children are linked by index for the final flatten.
struct BuildNode {
    Bounds3 bounds;
    int child[2] = {-1, -1};
    int axis = 0, firstPrim = 0, nPrims = 0;
};

struct Builder {
    std::vector<PrimInfo> info;
    std::vector<BuildNode> tree;
    std::vector<uint32_t> ordered;
    int maxPrims;
    int method;

    int leaf(int node, int start, int end, const Bounds3 &b) {
        tree[node].firstPrim = (int)ordered.size();
        tree[node].nPrims = end - start;
        tree[node].bounds = b;
        for (int i = start; i < end; ++i) ordered.push_back((uint32_t)info[i].number);
        return node;
    }

    // Decide the split of [start,end).  Returns -1 for "make a leaf".
    int split(int start, int end, const Bounds3 &bounds, int *dimOut) {
        int n = end - start;
        Bounds3 cb;
        for (int i = start; i < end; ++i) cb = unite(cb, info[i].centroid);
        int dim = cb.maximumExtent();
        *dimOut = dim;
        if (cb.pMax[dim] == cb.pMin[dim]) return -1;
        int mid = (start + end) / 2;
        auto byCentroid = [dim](const PrimInfo &a, const PrimInfo &b) {
            return a.centroid[dim] < b.centroid[dim];
        };
        if (method == IILE_SPLIT_MIDDLE) {
            float pmid = (cb.pMin[dim] + cb.pMax[dim]) / 2;
            PrimInfo *m = std::partition(&info[start], &info[end - 1] + 1,
                                         [dim, pmid](const PrimInfo &pi) { return pi.centroid[dim] < pmid; });
            mid = (int)(m - &info[0]);
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
        for (int i = start; i < end; ++i) {
            int b = bucketOf(info[i]);
            buckets[b].count++;
            buckets[b].bounds = unite(buckets[b].bounds, info[i].bounds);
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
            PrimInfo *m = std::partition(&info[start], &info[end - 1] + 1,
                                         [&](const PrimInfo &pi) { return bucketOf(pi) <= minBucket; });
            return (int)(m - &info[0]);
        }
        return -1;
    }

    // Depth-first build with an explicit stack.  Each frame is visited twice:
    // first to split and schedule its children (second child first, see the
    // header comment), then to merge the children's bounds.
    struct Frame {
        int node, start, end, stage;
    };
    int build(int nPrims) {
        tree.reserve(2 * (size_t)nPrims);
        tree.emplace_back();
        std::vector<Frame> stack;
        stack.push_back({0, 0, nPrims, 0});
        while (!stack.empty()) {
            Frame fr = stack.back();
            stack.pop_back();
            if (fr.stage == 1) {
                BuildNode &nd = tree[fr.node];
                nd.bounds = unite(tree[nd.child[0]].bounds, tree[nd.child[1]].bounds);
                continue;
            }
            Bounds3 bounds;
            for (int i = fr.start; i < fr.end; ++i) bounds = unite(bounds, info[i].bounds);
            int n = fr.end - fr.start;
            if (n == 1) {
                leaf(fr.node, fr.start, fr.end, bounds);
                continue;
            }
            int dim = 0;
            int mid = split(fr.start, fr.end, bounds, &dim);
            if (mid < 0) {
                leaf(fr.node, fr.start, fr.end, bounds);
                continue;
            }
            if (mid == fr.start || mid == fr.end)
                throw std::runtime_error("BVH build: empty partition (the reference would fail its CHECK_NE here)");
            int c0 = (int)tree.size();
            tree.emplace_back();
            int c1 = (int)tree.size();
            tree.emplace_back();
            tree[fr.node].axis = dim;
            tree[fr.node].child[0] = c0;
            tree[fr.node].child[1] = c1;
            stack.push_back({fr.node, fr.start, fr.end, 1});
            // LIFO: push the first child, then the second, so the second is built first
            stack.push_back({c0, fr.start, mid, 0});
            stack.push_back({c1, mid, fr.end, 0});
        }
        return (int)tree.size();
    }
};

}  // namespace

void buildBVH(HostScene *scene) {
    auto t0 = std::chrono::steady_clock::now();
    HostScene &s = *scene;
    size_t n = s.prims.size();
    s.nodes.clear();
    s.primOrder.clear();
    if (n == 0) return;
    Builder b;
    b.maxPrims = std::min(255, s.setup.maxNodePrims);
    b.method = s.setup.splitMethod;
    b.info.resize(n);
    for (size_t i = 0; i < n; ++i) {
        Bounds3 wb = primWorldBound(s, s.prims[i]);
        b.info[i] = {i, wb, 0.5f * wb.pMin + 0.5f * wb.pMax};
    }
    b.ordered.reserve(n);
    int total = b.build((int)n);

    // flatten depth-first, first child adjacent (bvh.cpp:640-658)
    s.nodes.resize(total);
    int next = 0;
    struct Item { int build; int parentSlot; };
    std::vector<Item> stack;
    stack.push_back({0, -1});
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        const BuildNode &bn = b.tree[it.build];
        int slot = next++;
        if (it.parentSlot >= 0) s.nodes[it.parentSlot].offset = slot;  // this is a second child
        IileBVHNode &ln = s.nodes[slot];
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
            ln.axis = (uint8_t)bn.axis;
            ln.nPrims = 0;
            stack.push_back({bn.child[1], slot});
            stack.push_back({bn.child[0], -1});
        }
    }
    s.primOrder = std::move(b.ordered);
    s.worldBound = Bounds3(V3(s.nodes[0].bmin[0], s.nodes[0].bmin[1], s.nodes[0].bmin[2]),
                           V3(s.nodes[0].bmax[0], s.nodes[0].bmax[1], s.nodes[0].bmax[2]));
    s.buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace iile
