// pbrt_b200 — host C++ command line over the C ABI (include/iile_b200.h): the same
// .pbrt file in, the same PFM / EXR film out as the reference's `pbrt` binary
// (src/main/pbrt.cpp:96-215 for the option set, api.cpp:1632-1692 for what WorldEnd does),
// with the render running on the B200 through libiile_b200.so.
//
//   pbrt_b200 [--outfile <file>] [--quiet] [--device <n>] [--spp <n>] [--nthreads <n>] scene.pbrt...
//
// --nthreads is accepted and ignored (the host side of this path is single threaded);
// options that select other subsystems of the reference (--cat, --toply, --reference=,
// --iile*) are rejected rather than silently ignored.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

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
                 "  --spp <n>            Override the sampler's pixelsamples.\n"
                 "  --nthreads <num>     Accepted for compatibility; ignored.\n");
}

int main(int argc, char **argv) {
    std::vector<std::string> files;
    std::string outfile;
    bool quiet = false;
    int device = 0, spp = 0;
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
        if (const char *v = value("--nthreads", 10)) (void)v;
        else if (const char *v = value("--outfile", 9)) outfile = v;
        else if (const char *v = value("--device", 8)) device = std::atoi(v);
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
    if (iile_device_count() < 1) {
        std::fprintf(stderr, "pbrt_b200: no CUDA device available; this renderer has no CPU path\n");
        return 1;
    }
    if (iile_set_device(device) != IILE_OK) return fail("--device");
    for (const std::string &f : files) {
        auto t0 = std::chrono::steady_clock::now();
        IileScene *scene = nullptr;
        if (iile_scene_load_pbrt(f.c_str(), 1, &scene) != IILE_OK) return fail(f.c_str());
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
        if (iile_render(scene, &params, film.data(), &stats) != IILE_OK) return fail("render");
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
