// ref_harness mipmap — oracle for the host side of image textures (csrc/host/texture_host.cpp).
// TEST INFRASTRUCTURE: calls the UNMODIFIED reference's ImageTexture::GetTexture (src/textures/imagemap.cpp:53-105:
// ReadImage, flip, convertIn, MIPMap ctor) and dumps the pyramid it built, level by level, row-major; with a
// query file it also evaluates MIPMap::Lookup (src/core/mipmap.h:239-369) for every query.
//
//   ref_harness mipmap <image> <channels 1|3> <wrap repeat|black|clamp> <scale> <gamma 0|1> <out.bin>
//                      [queries.bin <lookups.bin> <trilinear 0|1> <maxaniso>]
// out.bin: int32 nLevels, then per level int32 w, h and w*h*channels float32.
// queries.bin: n x {s, t, ds0, dt0, ds1, dt1} float32; lookups.bin: n x channels float32.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "pbrt.h"
#include "api.h"
#include "memory.h"
#include "spectrum.h"
#include "texture.h"
#include "paramset.h"
#include "imageio.h"
#define private public
#include "mipmap.h"
#include "textures/imagemap.h"
#undef private

#include "harness.h"

using namespace pbrt;

template <typename T>
static void put(std::vector<float> *out, const T &v);
template <>
void put<Float>(std::vector<float> *out, const Float &v) { out->push_back(v); }
template <>
void put<RGBSpectrum>(std::vector<float> *out, const RGBSpectrum &v) {
    Float rgb[3];
    v.ToRGB(rgb);
    out->push_back(rgb[0]), out->push_back(rgb[1]), out->push_back(rgb[2]);
}

template <typename Tmem, typename Tret>
static int run(int argc, char **argv, ImageWrap wrap) {
    const float scale = (float)atof(argv[5]);
    const bool gamma = atoi(argv[6]) != 0;
    const bool tri = argc > 10 ? atoi(argv[10]) != 0 : false;
    const float maxAniso = argc > 11 ? (float)atof(argv[11]) : 8.f;
    MIPMap<Tmem> *mm = ImageTexture<Tmem, Tret>::GetTexture(argv[2], tri, maxAniso, wrap, scale, gamma);
    std::ofstream o(argv[7], std::ios::binary);
    const int32_t nLevels = mm->Levels();
    o.write((const char *)&nLevels, 4);
    for (int l = 0; l < nLevels; ++l) {
        const int32_t w = mm->pyramid[l]->uSize(), h = mm->pyramid[l]->vSize();
        o.write((const char *)&w, 4);
        o.write((const char *)&h, 4);
        std::vector<float> v;
        for (int t = 0; t < h; ++t)
            for (int s = 0; s < w; ++s) put<Tmem>(&v, (*mm->pyramid[l])(s, t));
        o.write((const char *)v.data(), v.size() * 4);
    }
    if (argc > 9) {
        std::ifstream q(argv[8], std::ios::binary | std::ios::ate);
        if (!q) die(std::string("cannot open ") + argv[8]);
        std::vector<float> qs((size_t)q.tellg() / 4);
        q.seekg(0);
        q.read((char *)qs.data(), qs.size() * 4);
        std::vector<float> v;
        for (size_t i = 0; i + 5 < qs.size(); i += 6)
            put<Tmem>(&v, mm->Lookup(Point2f(qs[i], qs[i + 1]), Vector2f(qs[i + 2], qs[i + 3]), Vector2f(qs[i + 4], qs[i + 5])));
        std::ofstream lo(argv[9], std::ios::binary);
        lo.write((const char *)v.data(), v.size() * 4);
    }
    printf("{\"levels\": %d, \"width\": %d, \"height\": %d}\n", nLevels, mm->Width(), mm->Height());
    return 0;
}

int cmdMipmap(int argc, char **argv) {
    if (argc < 8) die("usage: mipmap <image> <channels> <wrap> <scale> <gamma> <out.bin> [queries.bin lookups.bin trilinear maxaniso]");
    Options opt;
    opt.nThreads = 1;
    opt.quiet = true;
    pbrtInit(opt);
    const std::string w = argv[4];
    const ImageWrap wrap = w == "black" ? ImageWrap::Black : (w == "clamp" ? ImageWrap::Clamp : ImageWrap::Repeat);
    const int rc = atoi(argv[3]) == 1 ? run<Float, Float>(argc, argv, wrap) : run<RGBSpectrum, Spectrum>(argc, argv, wrap);
    pbrtCleanup();
    return rc;
}
