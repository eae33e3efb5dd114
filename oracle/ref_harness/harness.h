// oracle/ref_harness/harness.h — TEST INFRASTRUCTURE shared by the harness commands.
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "pbrt.h"
#include "accelerators/bvh.h"
#include "geometry.h"
#include "light.h"
#include "primitive.h"
#include "sampler.h"
#include "transform.h"

#include "../../include/iile_b200.h"

// the product's IILEFLT1 container (csrc/host/flat_io.cpp)
struct Flat {
    std::vector<float> P, N, UV;
    std::vector<IilePrim> prims;
    std::vector<IileSphere> spheres;
    std::vector<IileMaterial> materials;
    std::vector<IileLight> lights;
    IileRenderSetup setup;
};


// reference objects built from the flat scene
struct RefScene {
    std::vector<std::unique_ptr<pbrt::Transform>> xforms;
    std::vector<std::shared_ptr<pbrt::Primitive>> prims;  // scene order
    std::vector<std::shared_ptr<pbrt::Light>> lights;
    std::map<const pbrt::Primitive *, int> primIndex;
    std::shared_ptr<pbrt::BVHAccel> bvh;
};


void die(const std::string &m);
Flat loadFlat(const std::string &path);
void buildRefScene(const Flat &fs, RefScene &rs, bool deindex);
double now();
// one PCG32 stream per (pixel, sample): sequence = (sampleOffset + s) << 32 | pixel index inside sb
pbrt::Sampler *makePixelStreamSampler(int64_t spp, const pbrt::Bounds2i &sb, int sampleOffset);
// Film + camera of the scene's setup, as cmdRender builds them
std::shared_ptr<const pbrt::Camera> makeCamera(const IileRenderSetup &S, const std::string &out, pbrt::Film **filmOut);

// parsed.cpp: run `fn` on what the reference's parser built from a .pbrt file (inside pbrtWorldEnd; pbrtInit first)
namespace pbrt { struct IileHookArgs; }
void withParsedScene(const std::string &pbrtFile, const std::function<void(const pbrt::IileHookArgs &)> &fn);
bool isPbrtFile(const std::string &path);
// the camera (and its film) RenderOptions::MakeCamera would build from the parsed scene (api.cpp:1630-1665)
std::shared_ptr<const pbrt::Camera> parsedCamera(const pbrt::IileHookArgs &a, pbrt::Film **filmOut);
