// ref_harness — the reference's own parser in front of a harness command.
// TEST INFRASTRUCTURE.  The harness links the patched copy of api.cpp that oracle/ref_patch/patch_api.sh makes
// (pbrtWorldEnd offers RenderOptions' objects to iileWorldEndHook before MakeIntegrator / Render): here the hook
// hands the parsed lights and primitives — image textures, bump maps and all, built by the UNMODIFIED reference's
// api.cpp / MakeMaterial / MakeFloatTexture — to the command that asked for them, and the reference's own render
// is skipped.  The work has to happen INSIDE the hook: pbrtWorldEnd frees the MIPMaps right after
// (ImageTexture::ClearCache, api.cpp:1545-1546).
#include "harness.h"

#include "api.h"
#include "iile_hook.h"
#include "parser.h"
#include "camera.h"
#include "cameras/environment.h"
#include "cameras/orthographic.h"
#include "cameras/perspective.h"
#include "film.h"
#include "filters/box.h"
#include "filters/gaussian.h"
#include "filters/mitchell.h"
#include "filters/sinc.h"
#include "filters/triangle.h"

namespace {
const std::function<void(const pbrt::IileHookArgs &)> *g_atWorldEnd = nullptr;
}

namespace pbrt {
bool iileWorldEndHook(const IileHookArgs &args) {
    if (!g_atWorldEnd) return false;
    (*g_atWorldEnd)(args);
    return true;
}
}  // namespace pbrt

void withParsedScene(const std::string &pbrtFile, const std::function<void(const pbrt::IileHookArgs &)> &fn) {
    g_atWorldEnd = &fn;
    pbrt::ParseFile(pbrtFile);
    g_atWorldEnd = nullptr;
}

bool isPbrtFile(const std::string &path) {
    return path.size() > 5 && path.compare(path.size() - 5, 5, ".pbrt") == 0;
}

// RenderOptions::MakeCamera (api.cpp:1630-1665) through the reference's public Create* factories: MakeFilter /
// MakeFilm / MakeCamera themselves are local to api.cpp
std::shared_ptr<const pbrt::Camera> parsedCamera(const pbrt::IileHookArgs &a, pbrt::Film **filmOut) {
    using namespace pbrt;
    std::unique_ptr<Filter> filter;
    if (a.FilterName == "box") filter.reset(CreateBoxFilter(a.FilterParams));
    else if (a.FilterName == "gaussian") filter.reset(CreateGaussianFilter(a.FilterParams));
    else if (a.FilterName == "mitchell") filter.reset(CreateMitchellFilter(a.FilterParams));
    else if (a.FilterName == "sinc") filter.reset(CreateSincFilter(a.FilterParams));
    else if (a.FilterName == "triangle") filter.reset(CreateTriangleFilter(a.FilterParams));
    else die("parsedCamera: filter " + a.FilterName);
    Film *film = CreateFilm(a.FilmParams, std::move(filter));
    static Transform cam2world;  // AnimatedTransform keeps the pointers
    cam2world = a.CameraToWorld;
    AnimatedTransform animated(&cam2world, 0, &cam2world, 1);
    Camera *camera = nullptr;
    if (a.CameraName == "perspective") camera = CreatePerspectiveCamera(a.CameraParams, animated, film, nullptr);
    else if (a.CameraName == "orthographic") camera = CreateOrthographicCamera(a.CameraParams, animated, film, nullptr);
    else if (a.CameraName == "environment") camera = CreateEnvironmentCamera(a.CameraParams, animated, film, nullptr);
    else die("parsedCamera: camera " + a.CameraName);
    *filmOut = film;
    return std::shared_ptr<const Camera>(camera);
}
