// oracle/ref_patch/iile_hook.h — the seam the patched pbrtWorldEnd() calls ("if the maintainer
// wants to keep the reference's parser").  TEST INFRASTRUCTURE: demonstrates the drop-in inside the reference.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "pbrt.h"
#include "paramset.h"
#include "transform.h"

namespace pbrt {

// what RenderOptions (a struct local to api.cpp, :159-187) holds when WorldEnd is reached
struct IileHookArgs {
    const std::string &FilterName;
    const ParamSet &FilterParams;
    const std::string &FilmName;
    const ParamSet &FilmParams;
    const std::string &SamplerName;
    const ParamSet &SamplerParams;
    const std::string &AcceleratorName;
    const ParamSet &AcceleratorParams;
    const std::string &IntegratorName;
    const ParamSet &IntegratorParams;
    const std::string &CameraName;
    const ParamSet &CameraParams;
    const Transform &CameraToWorld;
    const std::vector<std::shared_ptr<Light>> &lights;
    const std::vector<std::shared_ptr<Primitive>> &primitives;
};

// true: the B200 library rendered the scene and wrote the image; false (with the reason on stderr): the
// scene uses something the hook does not flatten and the reference's own CPU path must run
bool iileWorldEndHook(const IileHookArgs &args);

}  // namespace pbrt
