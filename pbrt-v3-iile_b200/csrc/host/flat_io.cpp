// Flat-scene container: the IileSceneDesc arrays, serialised.  This is the
// hand-off format between the front-end, the C ABI (iile_scene_save_flat /
// iile_scene_load_flat) and the test harness that rebuilds the same scene out
// of the reference's own classes.  Layout (little endian):
//   char magic[8] = "IILEFLT1"; u32 counts[8] = {nVerts, hasN, hasUV, nPrims,
//   nSpheres, nMaterials, nLights, nNodes}; IileRenderSetup; then the arrays
//   P, N?, UV?, prims, spheres, materials, lights, nodes, primOrder (if nNodes).
// IILEFLT3 (scenes with textures, ABI 4) appends u64 counts[4] = {nTextures, nImages, nTexels,
//   nMaterialTextures} and the arrays textures, images, texels, materialTextures.
#include <cmath>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "host_scene.h"

namespace iile {

namespace {
struct File {
    FILE *f;
    File(const std::string &p, const char *mode) : f(std::fopen(p.c_str(), mode)) {
        if (!f) throw std::runtime_error(p + ": cannot open");
    }
    ~File() { if (f) std::fclose(f); }
    void w(const void *d, size_t n) {
        if (n && std::fwrite(d, 1, n, f) != n) throw std::runtime_error("flat scene: short write");
    }
    void r(void *d, size_t n) {
        if (n && std::fread(d, 1, n, f) != n) throw std::runtime_error("flat scene: short read");
    }
};
// IILEFLT2 = IILEFLT1 with 96-byte materials and 272-byte lights (ABI 2); IILEFLT1 files are still read
const char kMagic[8] = {'I', 'I', 'L', 'E', 'F', 'L', 'T', '2'};
const char kMagicV1[8] = {'I', 'I', 'L', 'E', 'F', 'L', 'T', '1'};
// IILEFLT3 = IILEFLT2 + the texture section; written only for scenes that have textures, so that readers of
// IILEFLT2 (the oracle's harness, which has no textures) refuse such a file instead of rendering it wrongly
const char kMagicV3[8] = {'I', 'I', 'L', 'E', 'F', 'L', 'T', '3'};
constexpr size_t kMaterialBytesV1 = 64, kLightBytesV1 = 40;
}  // namespace

void saveFlat(const HostScene &s, const std::string &path) {
    File f(path, "wb");
    const bool textured = !s.textures.empty() || !s.materialTextures.empty();
    f.w(textured ? kMagicV3 : kMagic, 8);
    uint32_t c[8] = {s.nVerts(), (uint32_t)!s.N.empty(), (uint32_t)!s.UV.empty(),
                     (uint32_t)s.prims.size(), (uint32_t)s.spheres.size(),
                     (uint32_t)s.materials.size(), (uint32_t)s.lights.size(),
                     (uint32_t)s.nodes.size()};
    f.w(c, sizeof(c));
    f.w(&s.setup, sizeof(s.setup));
    f.w(s.P.data(), s.P.size() * 4);
    f.w(s.N.data(), s.N.size() * 4);
    f.w(s.UV.data(), s.UV.size() * 4);
    f.w(s.prims.data(), s.prims.size() * sizeof(IilePrim));
    f.w(s.spheres.data(), s.spheres.size() * sizeof(IileSphere));
    f.w(s.materials.data(), s.materials.size() * sizeof(IileMaterial));
    f.w(s.lights.data(), s.lights.size() * sizeof(IileLight));
    f.w(s.nodes.data(), s.nodes.size() * sizeof(IileBVHNode));
    if (!s.nodes.empty()) f.w(s.primOrder.data(), s.primOrder.size() * 4);
    if (textured) {
        const uint64_t t[4] = {s.textures.size(), s.images.size(), s.texels.size(), s.materialTextures.size()};
        f.w(t, sizeof(t));
        f.w(s.textures.data(), s.textures.size() * sizeof(IileTexture));
        f.w(s.images.data(), s.images.size() * sizeof(IileImage));
        f.w(s.texels.data(), s.texels.size() * sizeof(float));
        f.w(s.materialTextures.data(), s.materialTextures.size() * sizeof(IileMaterialTextures));
    }
}

// A caller-supplied LinearBVHNode array (flat file or IileSceneDesc.nodes) is walked by the upload
// (buildWideNodes) and by every traversal, so its topology is checked here: depth-first layout as
// flattenBVHTree writes it (bvh.cpp:640-658) — an interior node's first child is the next node and its
// second child lies further on; every node is reached exactly once; leaves cover the primitive slots
// exactly once; the depth stays inside the reference's own nodesToVisit[64] (bvh.cpp:670).
static void validateNodes(const HostScene &s) {
    const size_t n = s.nodes.size(), nPrims = s.prims.size();
    if (n == 0) return;
    std::vector<uint8_t> seenNode(n, 0), seenSlot(nPrims, 0);
    struct Item { uint32_t node; int depth; };
    std::vector<Item> todo;
    todo.push_back({0, 1});
    size_t visited = 0, covered = 0;
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        if (it.node >= n) throw std::runtime_error("scene: BVH child index out of range");
        if (seenNode[it.node]) throw std::runtime_error("scene: BVH node reached twice");
        if (it.depth > 64) throw std::runtime_error("scene: BVH deeper than 64 levels");
        seenNode[it.node] = 1;
        ++visited;
        const IileBVHNode &N = s.nodes[it.node];
        for (int a = 0; a < 3; ++a)
            if (!(N.bmin[a] <= N.bmax[a])) throw std::runtime_error("scene: BVH node with an empty or NaN box");
        if (N.nPrims > 0) {
            if (N.offset < 0 || (size_t)N.offset + N.nPrims > nPrims)
                throw std::runtime_error("scene: BVH leaf range outside the primitive array");
            for (size_t k = (size_t)N.offset; k < (size_t)N.offset + N.nPrims; ++k) {
                if (seenSlot[k]) throw std::runtime_error("scene: primitive slot in two BVH leaves");
                seenSlot[k] = 1;
            }
            covered += N.nPrims;
        } else {
            if (N.axis > 2) throw std::runtime_error("scene: BVH interior node with split axis > 2");
            if (N.offset <= (int64_t)it.node + 1 || (size_t)N.offset >= n)
                throw std::runtime_error("scene: BVH second-child offset out of order");
            todo.push_back({(uint32_t)N.offset, it.depth + 1});
            todo.push_back({it.node + 1, it.depth + 1});
        }
    }
    if (visited != n) throw std::runtime_error("scene: BVH array has unreachable nodes");
    if (covered != nPrims) throw std::runtime_error("scene: BVH leaves do not cover every primitive");
}

// Texture records, pyramids and material slots are indexed on the device without further checks.
static void validateTextures(const HostScene &s) {
    for (const IileImage &im : s.images) {
        if (im.width <= 0 || im.height <= 0 || (im.width & (im.width - 1)) || (im.height & (im.height - 1)) ||
            (im.channels != 1 && im.channels != 3))
            throw std::runtime_error("scene: image pyramid with a bad size or channel count");
        int levels = 1;
        for (int m = std::max(im.width, im.height); m > 1; m >>= 1) ++levels;
        if (im.nLevels != levels) throw std::runtime_error("scene: image pyramid with a wrong level count");
        uint64_t need = 0;
        for (int l = 0; l < levels; ++l)
            need += (uint64_t)std::max(1, im.width >> l) * std::max(1, im.height >> l) * im.channels;
        if (im.texelOffset > s.texels.size() || need > s.texels.size() - im.texelOffset)
            throw std::runtime_error("scene: image pyramid outside the texel pool");
    }
    const int nTex = (int)s.textures.size();
    std::vector<int> depth(s.textures.size(), 0);  // operators above the leaves
    for (int k = 0; k < nTex; ++k) {
        const IileTexture &t = s.textures[k];
        if (t.channels != 1 && t.channels != 3) throw std::runtime_error("scene: texture with a bad channel count");
        if (t.type == IILE_TEX_IMAGEMAP) {
            if (t.image < 0 || (size_t)t.image >= s.images.size() || s.images[t.image].channels != t.channels)
                throw std::runtime_error("scene: imagemap texture without a matching image");
            if (t.wrap < IILE_WRAP_REPEAT || t.wrap > IILE_WRAP_CLAMP) throw std::runtime_error("scene: unknown texture wrap mode");
            if (!(t.maxAnisotropy >= 0.f && t.maxAnisotropy <= 64.f)) throw std::runtime_error("scene: imagemap maxanisotropy out of range");
        } else if (t.type == IILE_TEX_SCALE || t.type == IILE_TEX_MIX || t.type == IILE_TEX_CHECKERBOARD) {
            // operands come earlier in the array: no cycles, and the device evaluates them without a stack
            for (int o : {t.tex1, t.tex2}) {
                if (o < 0 || o >= k || s.textures[o].channels != t.channels)
                    throw std::runtime_error("scene: scale / mix texture with a bad operand");
                depth[k] = std::max(depth[k], depth[o] + 1);
            }
            if (t.type == IILE_TEX_MIX && t.image != -1) {
                if (t.image < 0 || t.image >= k || s.textures[t.image].channels != 1)
                    throw std::runtime_error("scene: mix texture with a bad amount texture");
                depth[k] = std::max(depth[k], depth[t.image] + 1);
            }
            if (depth[k] > IILE_TEX_MAX_DEPTH) throw std::runtime_error("scene: scale / mix textures nested too deep");
        } else if (t.type != IILE_TEX_CONSTANT)
            throw std::runtime_error("scene: unknown texture type");
    }
    for (const IileMaterialTextures &mt : s.materialTextures)
        for (int k = 0; k < IILE_TEXSLOT_COUNT; ++k) {
            if (mt.slot[k] < -1 || mt.slot[k] >= nTex) throw std::runtime_error("scene: material texture slot out of range");
            if (mt.slot[k] >= 0) {
                const bool spectrum = k <= IILE_TEXSLOT_OPACITY;
                if (s.textures[mt.slot[k]].channels != (spectrum ? 3 : 1))
                    throw std::runtime_error("scene: material texture slot with the wrong texture type");
            }
        }
    for (const IileMaterial &m : s.materials)
        if (m.texSet < 0 || (size_t)m.texSet > s.materialTextures.size())
            throw std::runtime_error("scene: material texture set out of range");
}

// What a render setup may hold, wherever it comes from (flat file, iile_scene_create, iile_scene_set_setup, the
// parser as a backstop): a resolution, crop window or filter radius out of range sizes films and filter tables, an
// unknown enum selects no kernel, a NaN matrix only wastes the render — all refused here instead.
void validateSetup(const IileRenderSetup &S) {
    auto finite = [](const float *v, int n) {
        for (int k = 0; k < n; ++k)
            if (!std::isfinite(v[k])) return false;
        return true;
    };
    auto need = [](bool ok, const char *what) {
        if (!ok) throw std::runtime_error(std::string("render setup: ") + what);
    };
    need(finite(S.cameraToWorld, 16) && finite(S.worldToCamera, 16), "camera matrices must be finite");
    need(finite(&S.fov, 1) && S.fov > 0 && S.fov < 180, "fov must lie in (0, 180)");
    need(finite(&S.lensRadius, 1) && S.lensRadius >= 0 && finite(&S.focalDistance, 1), "lens radius / focal distance");
    need(finite(S.screenWindow, 4) && finite(&S.shutterOpen, 1) && finite(&S.shutterClose, 1), "screen window / shutter must be finite");
    need(S.xRes >= 1 && S.yRes >= 1 && S.xRes <= 65536 && S.yRes <= 65536, "resolution must lie in [1, 65536]");
    need(finite(S.cropWindow, 4) && S.cropWindow[0] >= 0 && S.cropWindow[1] <= 1 && S.cropWindow[2] >= 0 && S.cropWindow[3] <= 1 &&
             S.cropWindow[0] <= S.cropWindow[1] && S.cropWindow[2] <= S.cropWindow[3], "crop window must be ordered inside [0, 1]");
    need(finite(&S.filmScale, 1) && finite(&S.diagonal, 1) && !std::isnan(S.maxSampleLuminance), "film scale / diagonal / maxsampleluminance");
    need(S.filterType >= IILE_FILTER_BOX && S.filterType <= IILE_FILTER_MONITOR, "unknown pixel filter");
    need(finite(S.filterRadius, 2) && S.filterRadius[0] > 0 && S.filterRadius[1] > 0 && S.filterRadius[0] <= 64 && S.filterRadius[1] <= 64,
         "filter radius must lie in (0, 64]");
    need(finite(S.filterParam, 2), "filter parameters must be finite");
    need(S.pixelSamples >= 1 && S.pixelSamples <= (1 << 24), "pixelsamples must lie in [1, 2^24]");
    need(S.maxDepth >= 0 && S.maxDepth <= 100000, "maxdepth out of range");
    need(!std::isnan(S.rrThreshold), "rrthreshold is NaN");
    need(S.lightSampleStrategy >= 0 && S.lightSampleStrategy <= 2, "unknown lightsamplestrategy");
    need(S.maxNodePrims >= 1, "maxnodeprims must be positive");
    need(S.splitMethod >= 0 && S.splitMethod <= 3, "unknown splitmethod");
    need(S.cameraType >= 0 && S.cameraType <= 2, "unknown camera type");
    need(S.integratorType >= 0 && S.integratorType <= 2, "unknown integrator type");
    need(S.directStrategy >= 0 && S.directStrategy <= 1, "unknown direct-lighting strategy");
    need(S.samplerType >= 0 && S.samplerType <= 2, "unknown sampler type");
}

static void validate(const HostScene &s) {
    validateSetup(s.setup);
    uint32_t nv = s.nVerts();
    for (const IilePrim &p : s.prims) {
        if (p.type == IILE_PRIM_TRIANGLE) {
            for (int k = 0; k < 3; ++k)
                if (p.v[k] < 0 || (uint32_t)p.v[k] >= nv) throw std::runtime_error("scene: triangle vertex index out of range");
            if ((p.flags & IILE_PRIM_HAS_N) && s.N.empty()) throw std::runtime_error("scene: HAS_N primitive but no N array");
            if ((p.flags & IILE_PRIM_HAS_UV) && s.UV.empty()) throw std::runtime_error("scene: HAS_UV primitive but no UV array");
        } else if (p.type == IILE_PRIM_SPHERE) {
            if (p.v[0] < 0 || (size_t)p.v[0] >= s.spheres.size()) throw std::runtime_error("scene: sphere index out of range");
        } else
            throw std::runtime_error("scene: unknown primitive type");
        if (p.material < -1 || p.material >= (int)s.materials.size()) throw std::runtime_error("scene: material index out of range");
        if (p.light < -1 || p.light >= (int)s.lights.size()) throw std::runtime_error("scene: light index out of range");
    }
    validateTextures(s);
    for (const IileLight &l : s.lights)
        if (l.type == IILE_LIGHT_DIFFUSE_AREA && (l.prim < 0 || (size_t)l.prim >= s.prims.size()))
            throw std::runtime_error("scene: area light primitive out of range");
    if (!s.nodes.empty() && s.primOrder.size() != s.prims.size())
        throw std::runtime_error("scene: primOrder must list every primitive");
    {
        std::vector<uint8_t> seen(s.prims.size(), 0);
        for (uint32_t o : s.primOrder) {
            if (o >= s.prims.size()) throw std::runtime_error("scene: primOrder entry out of range");
            if (seen[o]) throw std::runtime_error("scene: primOrder lists a primitive twice");
            seen[o] = 1;
        }
    }
    validateNodes(s);
}

void loadFlat(const std::string &path, HostScene *s) {
    File f(path, "rb");
    char magic[8];
    f.r(magic, 8);
    const bool v1 = !std::memcmp(magic, kMagicV1, 8), v3 = !std::memcmp(magic, kMagicV3, 8);
    if (!v1 && !v3 && std::memcmp(magic, kMagic, 8)) throw std::runtime_error(path + ": not a flat scene file");
    uint32_t c[8];
    f.r(c, sizeof(c));
    f.r(&s->setup, sizeof(s->setup));
    s->P.resize((size_t)c[0] * 3);
    f.r(s->P.data(), s->P.size() * 4);
    s->N.resize(c[1] ? (size_t)c[0] * 3 : 0);
    f.r(s->N.data(), s->N.size() * 4);
    s->UV.resize(c[2] ? (size_t)c[0] * 2 : 0);
    f.r(s->UV.data(), s->UV.size() * 4);
    s->prims.resize(c[3]);
    f.r(s->prims.data(), s->prims.size() * sizeof(IilePrim));
    s->spheres.resize(c[4]);
    f.r(s->spheres.data(), s->spheres.size() * sizeof(IileSphere));
    s->materials.resize(c[5]);
    if (v1) {
        for (IileMaterial &m : s->materials) {
            std::memset(&m, 0, sizeof(m));
            f.r(&m, kMaterialBytesV1);
        }
    } else
        f.r(s->materials.data(), s->materials.size() * sizeof(IileMaterial));
    s->lights.resize(c[6]);
    if (v1) {
        for (IileLight &l : s->lights) {
            std::memset(&l, 0, sizeof(l));
            f.r(&l, kLightBytesV1);
        }
    } else
        f.r(s->lights.data(), s->lights.size() * sizeof(IileLight));
    s->nodes.resize(c[7]);
    f.r(s->nodes.data(), s->nodes.size() * sizeof(IileBVHNode));
    s->primOrder.resize(c[7] ? c[3] : 0);
    f.r(s->primOrder.data(), s->primOrder.size() * 4);
    if (v3) {
        uint64_t t[4];
        f.r(t, sizeof(t));
        if (t[0] > (1u << 24) || t[1] > (1u << 24) || t[2] > ((uint64_t)1 << 34) || t[3] > (1u << 24))
            throw std::runtime_error(path + ": texture section out of range");
        s->textures.resize(t[0]);
        f.r(s->textures.data(), s->textures.size() * sizeof(IileTexture));
        s->images.resize(t[1]);
        f.r(s->images.data(), s->images.size() * sizeof(IileImage));
        s->texels.resize(t[2]);
        f.r(s->texels.data(), s->texels.size() * sizeof(float));
        s->materialTextures.resize(t[3]);
        f.r(s->materialTextures.data(), s->materialTextures.size() * sizeof(IileMaterialTextures));
    }
    validate(*s);
    if (s->nodes.empty()) buildBVH(s);
    else {
        s->worldBound = Bounds3(V3(s->nodes[0].bmin[0], s->nodes[0].bmin[1], s->nodes[0].bmin[2]),
                                V3(s->nodes[0].bmax[0], s->nodes[0].bmax[1], s->nodes[0].bmax[2]));
        setLightWorldRadius(s);
    }
}

void fromDesc(const IileSceneDesc &d, HostScene *s) {
    if (!d.P || !d.prims || d.nPrims == 0) throw std::runtime_error("scene description has no geometry");
    s->P.assign(d.P, d.P + (size_t)d.nVerts * 3);
    if (d.N) s->N.assign(d.N, d.N + (size_t)d.nVerts * 3);
    if (d.UV) s->UV.assign(d.UV, d.UV + (size_t)d.nVerts * 2);
    s->prims.assign(d.prims, d.prims + d.nPrims);
    if (d.nSpheres) s->spheres.assign(d.spheres, d.spheres + d.nSpheres);
    if (d.nMaterials) s->materials.assign(d.materials, d.materials + d.nMaterials);
    if (d.nLights) s->lights.assign(d.lights, d.lights + d.nLights);
    s->setup = d.setup;
    if (d.nTextures && d.textures) s->textures.assign(d.textures, d.textures + d.nTextures);
    if (d.nImages && d.images) s->images.assign(d.images, d.images + d.nImages);
    if (d.nTexels && d.texels) s->texels.assign(d.texels, d.texels + d.nTexels);
    if (d.nMaterialTextures && d.materialTextures)
        s->materialTextures.assign(d.materialTextures, d.materialTextures + d.nMaterialTextures);
    if (d.nNodes && d.nodes && d.primOrder) {
        s->nodes.assign(d.nodes, d.nodes + d.nNodes);
        s->primOrder.assign(d.primOrder, d.primOrder + d.nPrims);
    }
    validate(*s);
    if (s->nodes.empty()) buildBVH(s);
    else {
        s->worldBound = Bounds3(V3(s->nodes[0].bmin[0], s->nodes[0].bmin[1], s->nodes[0].bmin[2]),
                                V3(s->nodes[0].bmax[0], s->nodes[0].bmax[1], s->nodes[0].bmax[2]));
        setLightWorldRadius(s);
    }
}

}  // namespace iile
