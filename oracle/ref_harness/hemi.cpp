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

using namespace pbrt;

int cmdHemi(int argc, char **argv) {
    if (argc < 5) die("usage: hemi <scene.flat> <points.bin> <out_prefix> [size] [spp] [nthreads]");
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
    double t1 = now();
    printf("{\"hemispheres\": %zu, \"size\": %d, \"spp\": %d, \"seconds\": %.6f, \"hemispheres_per_s\": %.3f}\n", n,
           size, spp, t1 - t0, n / (t1 - t0));
    pbrtCleanup();
    return 0;
}
