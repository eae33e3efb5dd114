// HostScene: the flattened, SoA form of a parsed .pbrt scene — exactly the
// arrays IileSceneDesc (include/iile_b200.h) exposes across the C ABI.
#pragma once

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../../include/iile_b200.h"
#include "vecmath.h"

namespace iile {

struct HostScene {
    std::vector<float> P, N, UV;  // N / UV are either empty or as long as P's vertex count
    std::vector<IilePrim> prims;  // scene-description order
    std::vector<IileSphere> spheres;
    std::vector<IileMaterial> materials;
    std::vector<IileLight> lights;
    // textures (ABI 4): texture records, MIPMap pyramids in one texel pool, per-material texture slots
    std::vector<IileTexture> textures;
    std::vector<IileImage> images;
    std::vector<float> texels;
    std::vector<IileMaterialTextures> materialTextures;
    // ImageTexture::textures (imagemap.h:118): one pyramid per (file, channels, wrap, scale, gamma)
    std::map<std::tuple<std::string, int, int, float, bool>, int> imageCache;
    std::vector<IileBVHNode> nodes;
    std::vector<uint32_t> primOrder;  // BVH slot -> scene-order primitive
    IileRenderSetup setup;
    Bounds3 worldBound;
    double buildSeconds = 0;

    uint32_t nVerts() const { return (uint32_t)(P.size() / 3); }
    V3 vertex(int i) const { return V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]); }
    // grow the optional per-vertex pools so they cover every vertex
    void padVertexPools() {
        if (!N.empty()) N.resize(P.size(), 0.f);
        if (!UV.empty()) UV.resize(P.size() / 3 * 2, 0.f);
    }
};

void defaultRenderSetup(IileRenderSetup *s);

// Sobol' generator matrices linked into the library (sobol_tables.cpp)
struct SobolTables {
    int nDims, matrixSize, nVdc, nVdcInv;
    const uint32_t *matrices32;  // [nDims][matrixSize]
    const uint64_t *vdc;         // [nVdc][matrixSize], row m - 1 for a 2^m grid
    const uint64_t *vdcInv;      // [nVdcInv][matrixSize]
};
const SobolTables &sobolTables();

// scene_loader.cpp — ParseFile + the api.cpp directive state machine for the
// supported subset.  Throws std::runtime_error on anything unsupported.
void loadPbrtFile(const std::string &path, HostScene *scene);

// loopsubdiv.cpp — Shape "loopsubdiv" (loopsubdiv.cpp:149-400): returns limit
// positions, limit normals (unnormalised S x T) and triangle indices.
void loopSubdivide(int nLevels, const std::vector<int> &indices, const std::vector<V3> &P,
                   std::vector<V3> *outP, std::vector<V3> *outN, std::vector<int> *outIdx);

// bvh_build.cpp — BVHAccel ctor (bvh.cpp:183-402, 640-658).
Bounds3 primWorldBound(const HostScene &s, const IilePrim &p);
void buildBVH(HostScene *scene);
// flat_io.cpp: throws std::runtime_error naming the field that is out of range
void validateSetup(const IileRenderSetup &setup);
void setLightWorldRadius(HostScene *scene);
// splitmethod "hlbvh" with the treelets built on the GPU (csrc/cuda/hlbvh_device.cu).  The CUDA side of the
// library installs the hook; buildBVH tries it first for HLBVH scenes of at least 2^16 primitives when a
// device is present (IILE_DEVICE_HLBVH=0 keeps the host build) and falls back to the host build — the same
// bytes either way — when it reports false.
using DeviceHlbvhFn = bool (*)(HostScene *scene, std::string *why);
extern DeviceHlbvhFn g_deviceHlbvh;
// buildUpperSAH (bvh.cpp:514-638) over the treelet roots: roots[t] = {bounds, nodes in treelet t}.  Returns
// the upper tree in creation order; out[0 .. nTreelets) are the treelets themselves (child = -1), the rest
// interior nodes; *root is the index of the tree's root.
struct HlbvhUpperNode {
    Bounds3 bounds;
    int child[2];
    int axis;
    int count;  // nodes of the flattened subtree
};
void hlbvhUpperTree(const std::vector<Bounds3> &rootBounds, const std::vector<int> &rootCounts,
                    std::vector<HlbvhUpperNode> *out, int *root);

// texture_host.cpp — ReadImage + ImageTexture::GetTexture + the MIPMap pyramid; returns the index of the
// image in scene->images (cached per file and conversion)
int loadImageMap(HostScene *scene, const std::string &file, int channels, int wrap, float scale, bool gamma);

// flat_io.cpp
void saveFlat(const HostScene &s, const std::string &path);
void loadFlat(const std::string &path, HostScene *s);
void fromDesc(const IileSceneDesc &d, HostScene *s);

// Shape::Area for a primitive (triangle.cpp:546-552, sphere.cpp:217)
float primArea(const HostScene &s, const IilePrim &p);

// plymesh.cpp — Shape "plymesh" (binary/ascii PLY, triangles and quads)
void loadPly(const std::string &path, std::vector<V3> *P, std::vector<V3> *N, std::vector<float> *UV,
             std::vector<int> *idx);

}  // namespace iile
