// pbrt_b200 — host C++ command line over the C ABI (include/iile_b200.h): the same
// .pbrt file in, the same PFM / EXR film out as the reference's `pbrt` binary
// (src/main/pbrt.cpp:96-215 for the option set, api.cpp:1632-1692 for what WorldEnd does),
// with the render running on the B200 through libiile_b200.so.
//
//   pbrt_b200 [--outfile <file>] [--quiet] [--device <n>] [--gpus <n>] [--spp <n>] [--nthreads <n>]
//             [--flatten <file.flat>] scene.pbrt...
//
// --gpus n: one host thread per GPU (devices 0..n-1).  The scene is flattened and its BVH built
// once, every thread uploads the same flat scene to its own device and renders a sample range of
// every pixel (SURVEY §8e partition A); the {R,G,B,w} films are summed onto GPU 0 with one NCCL
// reduce over NVLink (iile_film_reduce_dev) — Film::MergeFilmTile (film.cpp:135-148) across GPUs.
//
// --nthreads bounds the host threads of the BVH build (default: every core; the tree does not
// depend on it); options that select other subsystems of the reference (--cat, --toply, --reference=,
// --iile*) are rejected rather than silently ignored.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../../include/iile_b200.h"

static int fail(const char *what) {
    std::fprintf(stderr, "pbrt_b200: %s: %s\n", what, iile_last_error());
    return 1;
}

static void usage(const char *msg) {
    if (msg) std::fprintf(stderr, "pbrt_b200: %s\n\n", msg);
    std::fprintf(stderr,
                 "usage: pbrt_b200 [<options>] <filename.pbrt...>\n"
                 "  --help               Print this help text.\n"
                 "  --outfile <filename> Write the final image to the given filename (.pfm / .exr).\n"
                 "  --quiet              Suppress all text output other than error messages.\n"
                 "  --device <n>         CUDA device to render on (default 0).\n"
                 "  --gpus <n>           Render on devices 0..n-1, sample range per GPU, NCCL film reduce.\n"
                 "  --spp <n>            Override the sampler's pixelsamples.\n"
                 "  --flatten <file>     Parse, flatten and build the BVH on the host only; write the flat scene.\n"
                 "  --nthreads <num>     Host threads for the BVH build (default: all cores).\n");
}

// One rank of the multi-GPU render: its own device, scene copy, film and NCCL rank.
struct RankJob {
    int rank = 0, world = 1, spp = 0;
    size_t nPix = 0;
    std::string flat;
    const uint8_t *commId = nullptr;
    std::vector<float> film;  // rank 0: the merged film
    IileRenderStats stats{};
    std::string error;
};

static void runRank(RankJob *job) {
    auto bail = [&](const char *what) {
        if (job->error.empty()) job->error = std::string(what) + ": " + iile_last_error();
    };
    IileScene *scene = nullptr;
    void *comm = nullptr, *film = nullptr;
    const size_t nPix = job->nPix;
    // Every rank must reach ncclCommInitRank and the reduce, or the others would wait forever: a
    // rank that cannot even get its communicator or its film ends the process; a rank whose scene
    // upload or render fails still joins the collective, with a zero film, and reports afterwards.
    if (iile_set_device(job->rank) != IILE_OK || iile_comm_create(job->commId, job->world, job->rank, &comm) != IILE_OK ||
        iile_dev_malloc(nPix * 16, &film) != IILE_OK) {
        std::fprintf(stderr, "pbrt_b200: GPU %d: device / NCCL / film setup failed: %s\n", job->rank, iile_last_error());
        std::_Exit(1);
    }
    bool rendered = false;
    if (iile_scene_load_flat(job->flat.c_str(), 1, &scene) != IILE_OK) bail("scene upload");
    if (scene) {
        // SURVEY §8e partition A: contiguous sample ranges, sizes differing by at most one
        const int base = job->spp / job->world, extra = job->spp % job->world;
        IileRenderParams p{};
        p.sampleBegin = job->rank * base + (job->rank < extra ? job->rank : extra);
        p.sampleEnd = p.sampleBegin + base + (job->rank < extra ? 1 : 0);
        p.clearFilm = 1;
        if (p.sampleEnd > p.sampleBegin) {
            if (iile_render_dev(scene, &p, (float *)film, &job->stats, nullptr) == IILE_OK) rendered = true;
            else bail("render");
        }
    }
    if (!rendered) {  // an empty sample range, or a failure above: contribute nothing
        std::vector<float> zero(nPix * 4, 0.f);
        if (iile_dev_upload(film, zero.data(), nPix * 16) != IILE_OK) bail("film clear");
    }
    if (iile_film_reduce_dev((float *)film, nPix, 0, comm, nullptr) != IILE_OK) bail("film reduce");
    if (iile_dev_synchronize(nullptr) != IILE_OK) bail("synchronize");
    if (job->rank == 0 && job->error.empty()) {
        job->film.resize(nPix * 4);
        if (iile_dev_download(job->film.data(), film, nPix * 16) != IILE_OK) bail("film read-back");
    }
    iile_dev_free(film);
    iile_comm_destroy(comm);
    if (scene) iile_scene_destroy(scene);
}

int main(int argc, char **argv) {
    std::vector<std::string> files;
    std::string outfile, flattenTo;
    bool quiet = false;
    int device = 0, spp = 0, gpus = 1;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        auto value = [&](const char *name, size_t n) -> const char * {
            if (!std::strncmp(a, name, n) && a[n] == '=') return a + n + 1;
            if (!std::strcmp(a, name) || !std::strcmp(a, name + 1)) {
                if (i + 1 == argc) {
                    usage((std::string("missing value after ") + name + " argument").c_str());
                    std::exit(1);
                }
                return argv[++i];
            }
            return nullptr;
        };
        if (const char *v = value("--nthreads", 10)) setenv("IILE_BUILD_THREADS", v, 1);  // host BVH build
        else if (const char *v = value("--outfile", 9)) outfile = v;
        else if (const char *v = value("--flatten", 9)) flattenTo = v;
        else if (const char *v = value("--device", 8)) device = std::atoi(v);
        else if (const char *v = value("--gpus", 6)) gpus = std::atoi(v);
        else if (const char *v = value("--spp", 5)) spp = std::atoi(v);
        else if (!std::strcmp(a, "--quiet") || !std::strcmp(a, "-quiet")) quiet = true;
        else if (!std::strcmp(a, "--help") || !std::strcmp(a, "-help") || !std::strcmp(a, "-h")) {
            usage(nullptr);
            return 0;
        } else if (a[0] == '-' && a[1] != '\0') {
            usage((std::string("option \"") + a + "\" is outside the B200 rendering path").c_str());
            return 1;
        } else
            files.push_back(a);
    }
    if (files.empty()) {
        usage("no scene file given (reading the scene from standard input is not supported)");
        return 1;
    }
    if (!flattenTo.empty()) {
        // host only: parse + flatten + BVH build, written as a flat scene file (no device needed)
        IileScene *scene = nullptr;
        if (iile_scene_load_pbrt(files[0].c_str(), 0, &scene) != IILE_OK) return fail(files[0].c_str());
        if (spp > 0) {
            IileSceneInfo info;
            if (iile_scene_info(scene, &info) != IILE_OK) return fail("scene info");
            info.setup.pixelSamples = spp;
            if (iile_scene_set_setup(scene, &info.setup) != IILE_OK) return fail("--spp");
        }
        if (iile_scene_save_flat(scene, flattenTo.c_str()) != IILE_OK) return fail(flattenTo.c_str());
        iile_scene_destroy(scene);
        return 0;
    }
    if (iile_device_count() < 1) {
        std::fprintf(stderr, "pbrt_b200: no CUDA device available; this renderer has no CPU path\n");
        return 1;
    }
    if (gpus < 1 || gpus > iile_device_count()) {
        std::fprintf(stderr, "pbrt_b200: --gpus %d, but %d CUDA device(s) are visible\n", gpus, iile_device_count());
        return 1;
    }
    if (iile_set_device(device) != IILE_OK) return fail("--device");
    for (const std::string &f : files) {
        auto t0 = std::chrono::steady_clock::now();
        IileScene *scene = nullptr;
        if (iile_scene_load_pbrt(f.c_str(), gpus > 1 ? 0 : 1, &scene) != IILE_OK) return fail(f.c_str());
        IileSceneInfo info;
        if (iile_scene_info(scene, &info) != IILE_OK) return fail("scene info");
        if (spp > 0) {
            info.setup.pixelSamples = spp;
            if (iile_scene_set_setup(scene, &info.setup) != IILE_OK) return fail("--spp");
        }
        auto t1 = std::chrono::steady_clock::now();
        const int w = info.croppedBounds[2] - info.croppedBounds[0], h = info.croppedBounds[3] - info.croppedBounds[1];
        std::vector<float> film((size_t)w * h * 4), rgb((size_t)w * h * 3);
        IileRenderParams params{};
        params.clearFilm = 1;
        IileRenderStats stats{};
        if (gpus > 1) {
            // flatten + build once, upload per GPU from the flat file
            const std::string flat = "/tmp/pbrt_b200_" + std::to_string((long)getpid()) + ".flat";
            if (iile_scene_save_flat(scene, flat.c_str()) != IILE_OK) return fail("flat scene");
            uint8_t commId[IILE_COMM_ID_BYTES];
            if (iile_comm_unique_id(commId) != IILE_OK) return fail("NCCL unique id");
            std::vector<RankJob> jobs(gpus);
            std::vector<std::thread> threads;
            for (int r = 0; r < gpus; ++r) {
                jobs[r].rank = r, jobs[r].world = gpus, jobs[r].flat = flat, jobs[r].commId = commId;
                jobs[r].spp = info.setup.pixelSamples, jobs[r].nPix = (size_t)w * h;
                threads.emplace_back(runRank, &jobs[r]);
            }
            for (auto &t : threads) t.join();
            std::remove(flat.c_str());
            for (const RankJob &j : jobs) {
                if (!j.error.empty()) {
                    std::fprintf(stderr, "pbrt_b200: GPU %d: %s\n", j.rank, j.error.c_str());
                    return 1;
                }
                stats.cameraRays += j.stats.cameraRays, stats.closestRays += j.stats.closestRays;
                stats.shadowRays += j.stats.shadowRays, stats.launches += j.stats.launches;
            }
            film = jobs[0].film;
        } else if (iile_render(scene, &params, film.data(), &stats) != IILE_OK)
            return fail("render");
        if (iile_film_resolve(scene, film.data(), rgb.data()) != IILE_OK) return fail("film resolve");
        // Film ctor: --outfile overrides the scene's "filename" (film.cpp:271-282)
        std::string name = !outfile.empty() ? outfile : std::string(info.setup.filename);
        if (iile_write_image(name.c_str(), rgb.data(), w, h, info.setup.xRes, info.setup.yRes,
                             info.croppedBounds[0], info.croppedBounds[1]) != IILE_OK)
            return fail(name.c_str());
        auto t2 = std::chrono::steady_clock::now();
        if (!quiet) {
            const double load = std::chrono::duration<double>(t1 - t0).count();
            const double total = std::chrono::duration<double>(t2 - t1).count();
            const double rays = (double)(stats.closestRays + stats.shadowRays);
            std::printf("%s: %u triangles, %u spheres, %u lights, %u BVH nodes (parse+build+upload %.2f s)\n",
                        f.c_str(), info.nTriangles, info.nSpheres, info.nLights, info.nNodes, load);
            if (gpus > 1)
                std::printf("Rendering on %d GPUs: %dx%d, %d spp, maxdepth %d: %.1f Msamples, %.1f Mrays, %llu kernel "
                            "launches, %.2f s with upload, NCCL film reduce, read-back and image write -> %s\n",
                            gpus, w, h, info.setup.pixelSamples, info.setup.maxDepth, stats.cameraRays / 1e6,
                            rays / 1e6, (unsigned long long)stats.launches, total, name.c_str());
            else
            std::printf("Rendering: %dx%d, %d spp, maxdepth %d: %.1f ms on device (%.1f Msamples/s, %.1f Mrays/s, "
                        "%llu kernel launches), %.2f s with read-back and image write -> %s\n",
                        w, h, info.setup.pixelSamples, info.setup.maxDepth, stats.deviceSeconds * 1e3,
                        stats.cameraRays / stats.deviceSeconds / 1e6, rays / stats.deviceSeconds / 1e6,
                        (unsigned long long)stats.launches, total, name.c_str());
        }
        iile_scene_destroy(scene);
    }
    return 0;
}
