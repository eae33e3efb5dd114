// HostScene: the flattened, SoA form of a parsed .pbrt scene — exactly the
// arrays IileSceneDesc (include/iile_b200.h) exposes across the C ABI.
#pragma once

#include <string>
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
void setLightWorldRadius(HostScene *scene);

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
