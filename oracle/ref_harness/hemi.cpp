// ref_harness hemi — oracle for BASELINE configs[3], the IISPT hemispheric auxiliary pass.
// TEST INFRASTRUCTURE: harness code driving the UNMODIFIED reference classes
// (HemisphericCamera via CreateHemisphericCamera, src/cameras/hemispheric.cpp:108-158;
// IISPTdIntegrator through its public constructor, src/integrators/iispt_d.h:83-108;
// RenderView, iispt_d.cpp:345-461) over a flat scene.
//
//   ref_harness hemi <scene.flat> <points.bin> <out_prefix> [size] [spp] [nthreads]
//
// points.bin: n x {pos[3], dir[3]} float32.  For hemisphere h the files
// <out_prefix>_d_<h>.pfm (intensity, Film::WriteImage), _n_<h>.pfm (camera-space normals) and
// _z_<h>.pfm (distance, -1 = miss) are written exactly as IisptRenderRunner / --reference mode
// writes them (iispt.cpp:456-526).  The sampler is the harness's PixelStreamSampler — one PCG32
// stream per (pixel, sample), sequence = (h*spp + s) << 32 | (y*size + x) — the streams the GPU
// batch draws from; the reference's own RandomSampler(1, seed) serialises a tile.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "pbrt.h"
#include "api.h"
#include "camera.h"
#include "cameras/hemispheric.h"
#include "film.h"
#include "integrators/iispt_d.h"
#include "parallel.h"
#include "sampler.h"
#include "scene.h"

#include "harness.h"
#include "iile_hook.h"

// hemi_nn calls the reference's own IisptRenderRunner::normalizeMapsDownstream /
// transformMapsUpstream (iisptrenderrunner.cpp:1041-1136), which the class keeps private: every
// header it pulls in is included normally first (include guards), then the class definition alone
// is read with its access specifier opened.  Harness-side only; the reference tree is untouched.
#include <climits>
#include <unordered_map>
#include "integrators/iispt.h"
#include "integrators/iisptfilmmonitor.h"
#include "integrators/iisptnnconnector.h"
#include "integrators/iisptschedulemonitor.h"
#include "integrators/directlighting.h"
#include "integrators/directprogressiveintegrator.h"
#include "rng.h"
#include "tools/iisptmathutils.h"
#include "tools/iisptrng.h"
#include "tools/iisptpoint2i.h"
#include "samplers/sobol.h"
#include "tools/threadpool.h"
#include "tools/generalutils.h"
#define private public
#include "integrators/iisptrenderrunner.h"
#undef private

using namespace pbrt;

// the floats IisptNnConnector::pipe_image_film writes for one film (iisptnnconnector.cpp:36-72)
static void pipeFilm(std::shared_ptr<ImageFilm> film, std::vector<float> *out) {
    const int h = film->get_height(), w = film->get_width(), c = film->get_components();
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            PfmItem px = film->get(x, y);
            if (c == 1) out->push_back(px.get_single_component());
            else {
                Float r, g, b;
                px.get_triple_component(r, g, b);
                out->push_back(r), out->push_back(g), out->push_back(b);
            }
        }
}

//   ref_harness hemi_nn <scene.flat> <points.bin> <out_prefix> [size] [spp] [nthreads]
// Per hemisphere h: RenderView as `hemi`, then the runner's normalizeMapsDownstream; writes
// <out_prefix>_nn_<h>.bin = the 7*size*size floats IisptNnConnector::communicate pipes to the
// network (intensity HWC, normals HWC, distance HW) followed by rmean gmean bmean, and
// <out_prefix>_up_<h>.bin = transformMapsUpstream applied to the normalised intensity taken as the
// network's answer (size*size*3 floats, HWC): the identity network, enough to pin the arithmetic.
int cmdHemiNN(int argc, char **argv) {
    if (argc < 5) die("usage: hemi_nn <scene.flat> <points.bin> <out_prefix> [size] [spp] [nthreads]");
    Flat fs = loadFlat(argv[2]);
    const std::string prefix = argv[4];
    const int size = argc > 5 ? atoi(argv[5]) : 32;
    const int spp = argc > 6 ? atoi(argv[6]) : 1;
    Options opt;
    opt.nThreads = argc > 7 ? atoi(argv[7]) : 0;
    opt.quiet = true;
    opt.iisptHemiSize = size;
    pbrtInit(opt);
    PbrtOptions.iisptHemiSize = size;
    std::ifstream fp(argv[3], std::ios::binary | std::ios::ate);
    if (!fp) die(std::string("cannot open ") + argv[3]);
    const size_t n = (size_t)fp.tellg() / (6 * sizeof(float));
    fp.seekg(0);
    std::vector<float> pts(6 * n);
    fp.read((char *)pts.data(), pts.size() * sizeof(float));

    RefScene rs;
    buildRefScene(fs, rs, false);
    Scene scene(rs.bvh, rs.lights);
    Bounds2i pixelBounds(Point2i(0, 0), Point2i(size, size));
    std::shared_ptr<Sampler> runnerSampler(makePixelStreamSampler(1, pixelBounds, 0));
    IisptRenderRunner runner(nullptr, nullptr, nullptr, nullptr, nullptr, runnerSampler, 0, pixelBounds, nullptr);
    double t0 = now();
    for (size_t h = 0; h < n; ++h) {
        const float *p = &pts[6 * h];
        std::shared_ptr<Camera> camera(CreateHemisphericCamera(size, size, nullptr, Point3f(p[0], p[1], p[2]),
                                                               Vector3f(p[3], p[4], p[5]), prefix + "_unused.pfm"));
        std::shared_ptr<Sampler> sampler(makePixelStreamSampler(spp, pixelBounds, (int)h * spp));
        IISPTdIntegrator integ(3, camera, sampler, pixelBounds);
        integ.Preprocess(scene);
        integ.RenderView(scene, camera.get());
        std::unique_ptr<IntensityFilm> intensity = integ.get_intensity_film(camera.get());
        NormalFilm *normals = integ.get_normal_film();
        DistanceFilm *distance = integ.get_distance_film();
        float rmean, gmean, bmean;
        runner.normalizeMapsDownstream(intensity.get(), normals, distance, rmean, gmean, bmean);
        std::vector<float> out;
        pipeFilm(intensity->get_image_film(), &out);
        pipeFilm(normals->get_image_film(), &out);
        pipeFilm(distance->get_image_film(), &out);
        out.push_back(rmean), out.push_back(gmean), out.push_back(bmean);
        std::ofstream o1(prefix + "_nn_" + std::to_string(h) + ".bin", std::ios::binary);
        o1.write((const char *)out.data(), out.size() * sizeof(float));
        runner.transformMapsUpstream(intensity.get(), rmean, gmean, bmean);
        std::vector<float> up;
        pipeFilm(intensity->get_image_film(), &up);
        std::ofstream o2(prefix + "_up_" + std::to_string(h) + ".bin", std::ios::binary);
        o2.write((const char *)up.data(), up.size() * sizeof(float));
    }
    double t1 = now();
    printf("{\"hemispheres\": %zu, \"size\": %d, \"spp\": %d, \"seconds\": %.6f}\n", n, size, spp, t1 - t0);
    pbrtCleanup();
    return 0;
}

int cmdHemi(int argc, char **argv) {
    if (argc < 5) die("usage: hemi <scene.flat | scene.pbrt> <points.bin> <out_prefix> [size] [spp] [nthreads]");
    const std::string sceneFile = argv[2];
    const std::string prefix = argv[4];
    const int size = argc > 5 ? atoi(argv[5]) : 32;
    const int spp = argc > 6 ? atoi(argv[6]) : 1;
    Options opt;
    opt.nThreads = argc > 7 ? atoi(argv[7]) : 0;
    opt.quiet = true;
    opt.iisptHemiSize = size;
    pbrtInit(opt);
    PbrtOptions.iisptHemiSize = size;
    std::ifstream fp(argv[3], std::ios::binary | std::ios::ate);
    if (!fp) die(std::string("cannot open ") + argv[3]);
    const size_t n = (size_t)fp.tellg() / (6 * sizeof(float));
    fp.seekg(0);
    std::vector<float> pts(6 * n);
    fp.read((char *)pts.data(), pts.size() * sizeof(float));

    double seconds = 0;
    auto renderAll = [&](const Scene &scene) {
        double t0 = now();
        for (size_t h = 0; h < n; ++h) {
            const float *p = &pts[6 * h];
            std::string dName = prefix + "_d_" + std::to_string(h) + ".pfm";
            std::shared_ptr<Camera> camera(CreateHemisphericCamera(size, size, nullptr, Point3f(p[0], p[1], p[2]),
                                                                   Vector3f(p[3], p[4], p[5]), dName));
            Bounds2i pixelBounds(Point2i(0, 0), Point2i(size, size));
            std::shared_ptr<Sampler> sampler(makePixelStreamSampler(spp, pixelBounds, (int)h * spp));
            IISPTdIntegrator integ(3 /* CreateIISPTdIntegrator hard-codes maxdepth 3, iispt_d.cpp:505 */, camera,
                                   sampler, pixelBounds);
            integ.Preprocess(scene);
            integ.RenderView(scene, camera.get());
            integ.save_reference(camera, prefix + "_z_" + std::to_string(h) + ".pfm",
                                 prefix + "_n_" + std::to_string(h) + ".pfm");
        }
        seconds = now() - t0;
    };
    if (isPbrtFile(sceneFile)) {
        // the scene as the reference's own parser builds it (textures, bump maps): parsed.cpp
        withParsedScene(sceneFile, [&](const IileHookArgs &a) {
            std::shared_ptr<Primitive> accel = CreateBVHAccelerator(a.primitives, a.AcceleratorParams);
            Scene scene(accel, a.lights);
            renderAll(scene);
        });
    } else {
        Flat fs = loadFlat(sceneFile);
        RefScene rs;
        buildRefScene(fs, rs, false);
        Scene scene(rs.bvh, rs.lights);
        renderAll(scene);
    }
    printf("{\"hemispheres\": %zu, \"size\": %d, \"spp\": %d, \"seconds\": %.6f, \"hemispheres_per_s\": %.3f}\n", n,
           size, spp, seconds, n / seconds);
    pbrtCleanup();
    return 0;
}
