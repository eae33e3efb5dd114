// ref_harness direct_pass — oracle for IILE's direct-illumination pass: the UNMODIFIED reference's
// DirectProgressiveIntegrator (src/integrators/directprogressiveintegrator.cpp:5-150: preprocess, Li,
// RenderOnePass) accumulating into its IisptFilmMonitor (iisptfilmmonitor.cpp:47-75, :158-203), as
// IisptRenderRunner::run_direct drives them (iisptrenderrunner.cpp:601-631).  TEST INFRASTRUCTURE.
//
//   ref_harness direct_pass <scene.flat> <out.bin> <passes>
//
// RenderOnePass takes ONE camera sample per pixel (it never calls StartNextSample); run_direct calls it
// iileDirectSamples times on a sampler whose serial RNG keeps running.  Here pass p gets a PixelStreamSampler
// with sample offset p — stream (p << 32 | pixel), arrays from (1 << 62 | p << 32 | pixel) — i.e. pass p draws what
// sample p of the pixel draws on the device (DESIGN.md "Sample streams").  out.bin: int32 w, h, then w * h * 3
// floats, the film monitor's to_intensity_film() (sum of RGB / sum of weights per pixel) over
// Film::GetSampleBounds(), row-major from the bounds' corner.
#include <cstdio>
#include <fstream>
#include <memory>
#include <vector>

#include "pbrt.h"
#include "api.h"
#include "camera.h"
#include "film.h"
#include "film/intensityfilm.h"
#include "integrators/directprogressiveintegrator.h"
#include "integrators/iisptfilmmonitor.h"
#include "scene.h"

#include "harness.h"

using namespace pbrt;

int cmdDirectPass(int argc, char **argv) {
    if (argc < 5) die("usage: direct_pass <scene.flat> <out.bin> <passes>");
    Flat fs = loadFlat(argv[2]);
    const int passes = atoi(argv[4]);
    Options opt;
    opt.nThreads = 1;
    opt.quiet = true;
    pbrtInit(opt);
    RefScene rs;
    buildRefScene(fs, rs, false);
    Scene scene(rs.bvh, rs.lights);
    Film *film = nullptr;
    std::shared_ptr<const Camera> camera = makeCamera(fs.setup, std::string(argv[3]) + "_unused.pfm", &film);
    const Bounds2i bounds = film->GetSampleBounds();  // what IISPTIntegrator gives its film monitors, iispt.cpp:366-376
    IisptFilmMonitor monitor(bounds);
    double t0 = now();
    for (int p = 0; p < passes; ++p) {
        std::unique_ptr<Sampler> sampler(makePixelStreamSampler(1, bounds, p));
        DirectProgressiveIntegrator integ(camera, std::move(sampler), monitor.get_film_bounds());
        integ.preprocess(scene);
        integ.RenderOnePass(scene, &monitor);
    }
    double t1 = now();
    std::shared_ptr<IntensityFilm> out = monitor.to_intensity_film();
    std::shared_ptr<ImageFilm> img = out->get_image_film();
    const int w = bounds.pMax.x - bounds.pMin.x, h = bounds.pMax.y - bounds.pMin.y;
    std::vector<float> v;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            Float r, g, b;
            img->get(x, y).get_triple_component(r, g, b);
            v.push_back(r), v.push_back(g), v.push_back(b);
        }
    std::ofstream o(argv[3], std::ios::binary);
    const int32_t dims[2] = {w, h};
    o.write((const char *)dims, sizeof(dims));
    o.write((const char *)v.data(), v.size() * sizeof(float));
    printf("{\"passes\": %d, \"width\": %d, \"height\": %d, \"x0\": %d, \"y0\": %d, \"seconds\": %.6f}\n", passes, w, h,
           bounds.pMin.x, bounds.pMin.y, t1 - t0);
    pbrtCleanup();
    return 0;
}
