// ref_harness iispt_task — oracle for one task of IisptRenderRunner::run
// (src/integrators/iisptrenderrunner.cpp:216-596).  TEST INFRASTRUCTURE: harness code that walks a task the way
// run() does, calling the UNMODIFIED reference's own members for every step that computes something —
// sampler_next_pixel, find_intersection, compute_fpixel_weights, sample_hemisphere (and through it the
// file-static estimate_direct), HemisphericCamera::get_light_sample_nn / getLightSampleNn,
// normalizeMapsDownstream / transformMapsUpstream, IISPTdIntegrator::RenderView.  What run() needs around
// them and a harness cannot have is replaced:
//   * the schedule monitor hands out tasks      -> the task comes from the command line
//   * IisptNnConnector pipes the maps to a child Python process -> the films come from a file (the device
//     path's own films, which isolates the gather), or from the identity network on the reference's
//     own RenderView (normalise, take the intensity as the answer, transform back)
//   * the film monitor accumulates the samples   -> the samples are written out per pixel
//   * ONE serial IisptRng per thread              -> reseeded per film pixel, IisptRng(rngSeed + j), the
//     stream the device's gather kernel draws from (a kernel cannot share a serial stream)
//   * the scene's Sampler, cloned per thread      -> the harness's PixelStreamSampler: sequence =
//     samplerSample << 32 | sampler_pixel_counter.x
//
//   ref_harness iispt_task <scene.flat> <out_prefix> x0 y0 x1 y1 tilesize hemiSize samplerCounter
//                          samplerSample rngSeed [films.bin | -] [spp of RenderView]
//
// Writes <out_prefix>_hemi.bin   (per hemi point: valid, pos[3], dir[3] as 7 float32)
//        <out_prefix>_films.bin  (per valid hemi point: size*size*3 float32 in ImageFilm order — what
//                                 set_nn_film receives)
//        <out_prefix>_pixels.bin (per film pixel, row-major: f_beta*L rgb and the sample weight: 0.5, or 0
//                                 where run() adds no sample)
#include <climits>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <unordered_map>
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
// the runner keeps the members run() is made of private: as in hemi.cpp, every header is included normally
// first, then the class definition alone is read with its access specifier opened
#define private public
#include "integrators/iisptrenderrunner.h"
#undef private

using namespace pbrt;

static void filmFloats(std::shared_ptr<ImageFilm> film, std::vector<float> *out) {
    const int h = film->get_height(), w = film->get_width();
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            Float r, g, b;
            film->get(x, y).get_triple_component(r, g, b);
            out->push_back(r), out->push_back(g), out->push_back(b);
        }
}

namespace {
struct TaskArgs {
    std::string prefix, filmsFile;
    IisptScheduleMonitorTask task;
    int size, samplerCounter, samplerSample, spp;
    uint64_t rngSeed;
};

// the two loops of IisptRenderRunner::run over one task, on a scene and main camera built either from the flat file
// (harness materials: constants only) or by the reference's own parser (textures, bump maps: parsed.cpp)
// pixelsReturn (optional): the task's [y1-y0][x1-x0][4] samples come back instead of being written to files;
// cameraSamples: how far the runner's pixel counter moved
void runTask(const Scene &scene, std::shared_ptr<const Camera> mainCamera, Film *film, const TaskArgs &A,
             std::vector<float> *pixelsReturn = nullptr, int *cameraSamples = nullptr) {
    const std::string &prefix = A.prefix, &filmsFile = A.filmsFile;
    const IisptScheduleMonitorTask &task = A.task;
    const int size = A.size, samplerCounter = A.samplerCounter, samplerSample = A.samplerSample, spp = A.spp;
    const uint64_t rngSeed = A.rngSeed;
    const Bounds2i hemiBounds(Point2i(0, 0), Point2i(size, size));
    // sampler_next_pixel() starts pixel (counter, 0): stream index = the counter
    std::shared_ptr<Sampler> runnerSampler(makePixelStreamSampler(1, Bounds2i(Point2i(0, 0), Point2i(INT_MAX, 1)), samplerSample));
    IisptRenderRunner runner(nullptr, nullptr, nullptr, mainCamera, nullptr, runnerSampler, 0, film->GetSampleBounds(), nullptr);
    runner.sampler_pixel_counter = Point2i(samplerCounter, 0);
    const Point3f mainCameraOrigin = mainCamera->getCameraWorldPosition();

    std::vector<float> givenFilms;
    if (filmsFile != "-") {
        std::ifstream fp(filmsFile, std::ios::binary | std::ios::ate);
        if (!fp) die("cannot open " + filmsFile);
        givenFilms.resize((size_t)fp.tellg() / sizeof(float));
        fp.seekg(0);
        fp.read((char *)givenFilms.data(), givenFilms.size() * sizeof(float));
    }

    MemoryArena arena;
    std::unordered_map<IisptPoint2i, std::shared_ptr<HemisphericCamera>> hemi_points;
    std::vector<float> hemiOut, filmsOut;
    size_t nValid = 0;
    double t0 = now();
    // ---- the hemi points (run :252-389) -----------------------------------------------------------------
    int tile_x = task.x0, tile_y = task.y0;
    while (1) {
        IisptPoint2i hemi_key;
        hemi_key.x = tile_x;
        hemi_key.y = tile_y;
        Point2i pixel(tile_x, tile_y);
        runner.sampler_next_pixel();
        CameraSample camera_sample = runner.sampler->GetCameraSample(pixel);
        RayDifferential r;
        mainCamera->GenerateRayDifferential(camera_sample, &r);
        r.ScaleDifferentials(1.0);
        SurfaceInteraction isect;
        Spectrum beta, background, area_out;
        RayDifferential ray;
        bool found = runner.find_intersection(r, scene, arena, &isect, &ray, &beta, &background, &area_out);
        if (!found || beta.y() <= 0.0) {
            hemi_points[hemi_key] = nullptr;
            hemiOut.insert(hemiOut.end(), {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f});
        } else {
            Normal3f surface_normal = isect.n;
            if (Dot(Vector3f(isect.n.x, isect.n.y, isect.n.z), Vector3f(ray.d.x, ray.d.y, ray.d.z)) > 0.0)
                surface_normal = Normal3f(-isect.n.x, -isect.n.y, -isect.n.z);
            Ray aux_ray = isect.SpawnRay(Vector3f(surface_normal));
            std::shared_ptr<HemisphericCamera> aux_camera(
                CreateHemisphericCamera(size, size, nullptr, aux_ray.o, aux_ray.d, std::string("/tmp/null")));
            hemiOut.insert(hemiOut.end(), {1.f, aux_ray.o.x, aux_ray.o.y, aux_ray.o.z, aux_ray.d.x, aux_ray.d.y, aux_ray.d.z});
            std::shared_ptr<IntensityFilm> nn_film;
            if (!givenFilms.empty()) {
                if ((nValid + 1) * (size_t)size * size * 3 > givenFilms.size()) die("films.bin holds too few films");
                nn_film.reset(new IntensityFilm(size, size));
                nn_film->populate_from_float_array(&givenFilms[nValid * (size_t)size * size * 3]);
            } else {
                // RenderView with the streams of the device batch: hemisphere h = the h-th VALID hemi point
                std::shared_ptr<Camera> cam = aux_camera;
                std::shared_ptr<Sampler> sampler(makePixelStreamSampler(spp, hemiBounds, (int)nValid * spp));
                IISPTdIntegrator integ(3, cam, sampler, hemiBounds);
                integ.Preprocess(scene);
                integ.RenderView(scene, aux_camera.get());
                std::unique_ptr<IntensityFilm> intensity = integ.get_intensity_film(aux_camera.get());
                float rmean, gmean, bmean;
                runner.normalizeMapsDownstream(intensity.get(), integ.get_normal_film(), integ.get_distance_film(), rmean,
                                               gmean, bmean);
                runner.transformMapsUpstream(intensity.get(), rmean, gmean, bmean);  // the identity network's answer
                nn_film = std::move(intensity);
            }
            filmFloats(nn_film->get_image_film(), &filmsOut);
            aux_camera->set_nn_film(nn_film);
            hemi_points[hemi_key] = aux_camera;
            ++nValid;
        }
        bool advance_tile_y = false;
        if (tile_x == task.x1 - 1) {
            tile_x = task.x0;
            advance_tile_y = true;
        } else
            tile_x = std::min(tile_x + task.tilesize, task.x1 - 1);
        if (advance_tile_y) {
            if (tile_y == task.y1 - 1) break;
            tile_y = std::min(tile_y + task.tilesize, task.y1 - 1);
        }
    }
    double t1 = now();
    // ---- the film pixels (run :404-570) -----------------------------------------------------------------
    std::vector<float> pixelsOut;
    size_t nSamples = 0;
    for (int fy = task.y0; fy < task.y1; fy++) {
        for (int fx = task.x0; fx < task.x1; fx++) {
            const size_t j = (size_t)(fy - task.y0) * (task.x1 - task.x0) + (fx - task.x0);
            Point2i f_pixel(fx, fy);
            Point2i neigh_s(fx - (iispt::positiveModulo(fx - task.x0, task.tilesize)),
                            fy - (iispt::positiveModulo(fy - task.y0, task.tilesize)));
            Point2i neigh_e(std::min(neigh_s.x + task.tilesize, task.x1 - 1), std::min(neigh_s.y + task.tilesize, task.y1 - 1));
            Point2i neigh_r(neigh_e.x, neigh_s.y);
            Point2i neigh_b(neigh_s.x, neigh_e.y);
            Point2i neighbour_points[4] = {neigh_s, neigh_e, neigh_r, neigh_b};
            HemisphericCamera *hemi_sampling_cameras[4];
            for (int i = 0; i < 4; i++) {
                IisptPoint2i key;
                key.x = neighbour_points[i].x;
                key.y = neighbour_points[i].y;
                if (hemi_points.count(key) <= 0) die("hemi_points does not have a neighbour's key");
                hemi_sampling_cameras[i] = hemi_points.at(key).get();
            }
            runner.sampler_next_pixel();
            CameraSample f_camera_sample = runner.sampler->GetCameraSample(f_pixel);
            RayDifferential f_r;
            mainCamera->GenerateRayDifferential(f_camera_sample, &f_r);
            f_r.ScaleDifferentials(1.0);
            SurfaceInteraction f_isect;
            Spectrum f_beta, f_background, area_out;
            RayDifferential f_ray;
            bool f_found = runner.find_intersection(f_r, scene, arena, &f_isect, &f_ray, &f_beta, &f_background, &area_out);
            if (!f_found || f_beta.y() <= 0.0) {
                pixelsOut.insert(pixelsOut.end(), {0.f, 0.f, 0.f, 0.f});
                continue;
            }
            float hemi_sampling_weights[4];
            runner.compute_fpixel_weights(4, neighbour_points, hemi_sampling_cameras, f_pixel, f_isect, task.tilesize, f_ray,
                                          hemi_sampling_weights, mainCameraOrigin);
            f_isect.ComputeScatteringFunctions(f_ray, arena);
            if (!f_isect.bsdf) {
                pixelsOut.insert(pixelsOut.end(), {0.f, 0.f, 0.f, 0.f});
                continue;
            }
            runner.rng.reset(new IisptRng(rngSeed + j));  // see the header: one stream per film pixel
            Spectrum L(0.0);
            L += runner.sample_hemisphere(f_isect, 4, hemi_sampling_weights, hemi_sampling_cameras);
            Float rgb[3];
            Spectrum(f_beta * L).ToRGB(rgb);
            pixelsOut.insert(pixelsOut.end(), {rgb[0], rgb[1], rgb[2], 0.5f});
            ++nSamples;
            arena.Reset();
        }
    }
    double t2 = now();
    if (cameraSamples) *cameraSamples = runner.sampler_pixel_counter.x - samplerCounter;
    if (pixelsReturn) {
        *pixelsReturn = pixelsOut;
        return;
    }
    auto dump = [&](const std::string &name, const std::vector<float> &v) {
        std::ofstream o(prefix + name, std::ios::binary);
        o.write((const char *)v.data(), v.size() * sizeof(float));
    };
    dump("_hemi.bin", hemiOut);
    dump("_films.bin", filmsOut);
    dump("_pixels.bin", pixelsOut);
    printf("{\"hemi_points\": %zu, \"valid\": %zu, \"pixels\": %zu, \"samples\": %zu, \"hemi_seconds\": %.6f, "
           "\"gather_seconds\": %.6f}\n",
           hemiOut.size() / 7, nValid, pixelsOut.size() / 4, nSamples, t1 - t0, t2 - t1);
}
}  // namespace

int cmdIisptTask(int argc, char **argv) {
    if (argc < 13)
        die("usage: iispt_task <scene.flat | scene.pbrt> <out_prefix> x0 y0 x1 y1 tilesize hemiSize samplerCounter "
            "samplerSample rngSeed [films.bin | -] [spp]");
    const std::string sceneFile = argv[2];
    TaskArgs A;
    A.prefix = argv[3];
    IisptScheduleMonitorTask &task = A.task;
    task.x0 = atoi(argv[4]), task.y0 = atoi(argv[5]), task.x1 = atoi(argv[6]), task.y1 = atoi(argv[7]);
    task.tilesize = atoi(argv[8]);
    task.taskNumber = 0;
    const int size = A.size = atoi(argv[9]);
    A.samplerCounter = atoi(argv[10]), A.samplerSample = atoi(argv[11]);
    A.rngSeed = strtoull(argv[12], nullptr, 10);
    A.filmsFile = argc > 13 ? argv[13] : "-";
    A.spp = argc > 14 ? atoi(argv[14]) : 1;
    Options opt;
    opt.nThreads = 1;
    opt.quiet = true;
    opt.iisptHemiSize = size;
    pbrtInit(opt);
    PbrtOptions.iisptHemiSize = size;

    if (isPbrtFile(sceneFile)) {
        withParsedScene(sceneFile, [&](const IileHookArgs &a) {
            std::shared_ptr<Primitive> accel = CreateBVHAccelerator(a.primitives, a.AcceleratorParams);
            Scene scene(accel, a.lights);
            Film *film = nullptr;
            std::shared_ptr<const Camera> mainCamera = parsedCamera(a, &film);
            runTask(scene, mainCamera, film, A);
        });
    } else {
        Flat fs = loadFlat(sceneFile);
        RefScene rs;
        buildRefScene(fs, rs, false);
        Scene scene(rs.bvh, rs.lights);
        Film *film = nullptr;
        std::shared_ptr<const Camera> mainCamera = makeCamera(fs.setup, A.prefix + "_unused.pfm", &film);
        runTask(scene, mainCamera, film, A);
    }
    pbrtCleanup();
    return 0;
}

// ref_harness iile_render — the whole of IISPTIntegrator::render_normal_2 (iispt.cpp:357-447) on one thread, from the
// reference's own parts: IisptScheduleMonitor::next_task hands out the tasks (iisptschedulemonitor.cpp:41-76), each
// task runs as above (run's two loops, identity network on the reference's own RenderView), its samples go to
// film_monitor_indirect->add_n_samples; DirectProgressiveIntegrator::RenderOnePass fills film_monitor_direct
// (run_direct, iisptrenderrunner.cpp:601-631); merge_into adds the two normalised films (iisptfilmmonitor.cpp:231-275).
//
//   ref_harness iile_render <scene.flat> <out.bin> indirectTasks directSamples hemiSize radiusStart seed
//
// Streams: the runner's pixel counter runs on from task to task (one runner); task k seeds the per-pixel gather
// streams with seed + (k << 32); direct pass p draws the streams of sample p (direct_pass.cpp).  out.bin: int32 w, h,
// then the merged film, w * h * 3 floats over Film::GetSampleBounds().
#include "film/intensityfilm.h"
#include "integrators/directprogressiveintegrator.h"

int cmdIileRender(int argc, char **argv) {
    if (argc < 9) die("usage: iile_render <scene.flat> <out.bin> indirectTasks directSamples hemiSize radiusStart seed");
    Flat fs = loadFlat(argv[2]);
    const int indirectTasks = atoi(argv[4]), directSamples = atoi(argv[5]), size = atoi(argv[6]);
    setenv("IISPT_SCHEDULE_RADIUS_START", argv[7], 1);
    const uint64_t seed = strtoull(argv[8], nullptr, 10);
    Options opt;
    opt.nThreads = 1;
    opt.quiet = true;
    opt.iisptHemiSize = size;
    pbrtInit(opt);
    PbrtOptions.iisptHemiSize = size;
    RefScene rs;
    buildRefScene(fs, rs, false);
    Scene scene(rs.bvh, rs.lights);
    Film *film = nullptr;
    std::shared_ptr<const Camera> camera = makeCamera(fs.setup, std::string(argv[3]) + "_unused.pfm", &film);
    const Bounds2i bounds = film->GetSampleBounds();
    IisptScheduleMonitor schedule(bounds);
    IisptFilmMonitor indirect(bounds), direct(bounds);
    TaskArgs A;
    A.prefix = std::string(argv[3]) + "_task";
    A.filmsFile = "-";
    A.size = size;
    A.samplerSample = 0;
    A.spp = 1;
    int counter = 0, nTasks = 0;
    double t0 = now();
    for (;;) {
        IisptScheduleMonitorTask task = schedule.next_task();
        if (task.taskNumber >= indirectTasks) break;  // run :234-236
        A.task = task;
        A.samplerCounter = counter;
        A.rngSeed = seed + ((uint64_t)task.taskNumber << 32);
        std::vector<float> px;
        int used = 0;
        runTask(scene, camera, film, A, &px, &used);
        counter += used;
        std::vector<Point2i> pts;
        std::vector<Spectrum> ss;
        std::vector<double> ws;
        const int w = task.x1 - task.x0;
        for (size_t j = 0; j < px.size() / 4; ++j)
            if (px[4 * j + 3] > 0.f) {
                pts.push_back(Point2i(task.x0 + (int)(j % w), task.y0 + (int)(j / w)));
                Float rgb[3] = {px[4 * j], px[4 * j + 1], px[4 * j + 2]};
                ss.push_back(Spectrum::FromRGB(rgb));
                ws.push_back(px[4 * j + 3]);
            }
        indirect.add_n_samples(pts, ss, ws);
        ++nTasks;
    }
    double t1 = now();
    for (int p = 0; p < directSamples; ++p) {
        std::unique_ptr<Sampler> sampler(makePixelStreamSampler(1, bounds, p));
        DirectProgressiveIntegrator integ(camera, std::move(sampler), direct.get_film_bounds());
        integ.preprocess(scene);
        integ.RenderOnePass(scene, &direct);
    }
    double t2 = now();
    std::shared_ptr<IisptFilmMonitor> merged = direct.merge_into(&indirect);
    std::shared_ptr<ImageFilm> img = merged->to_intensity_film()->get_image_film();
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
    printf("{\"tasks\": %d, \"camera_samples\": %d, \"direct_passes\": %d, \"width\": %d, \"height\": %d, "
           "\"indirect_seconds\": %.6f, \"direct_seconds\": %.6f}\n",
           nTasks, counter, directSamples, w, h, t1 - t0, t2 - t1);
    pbrtCleanup();
    return 0;
}
