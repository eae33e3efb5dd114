// Film geometry, pixel-filter table, final pixel arithmetic and image files.
//
// Restates, on the host: Film::Film's cropped bounds and 16x16 filter table
// (src/core/film.cpp:45-75), GetSampleBounds (:77-83), the five pixel filters'
// Evaluate (src/filters/{box,gaussian,mitchell,sinc,triangle}.{h,cpp}),
// MergeFilmTile's RGB->XYZ (:135-148, spectrum.h:62-66), to_rgb_array's
// XYZ->RGB / weight / clamp / scale (:187-222), WriteImagePFM
// (src/core/imageio.cpp:437-482) and WriteImageEXR's half RGBA scan-line file
// with data / display windows (imageio.cpp:164-189; written here without
// OpenEXR as an uncompressed scan-line file, same pixels and windows).
#include "film_host.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace iile {

FilmGeometry filmGeometry(const IileRenderSetup &s) {
    FilmGeometry g;
    g.cropped[0] = (int)std::ceil(s.xRes * s.cropWindow[0]);
    g.cropped[1] = (int)std::ceil(s.yRes * s.cropWindow[2]);
    g.cropped[2] = (int)std::ceil(s.xRes * s.cropWindow[1]);
    g.cropped[3] = (int)std::ceil(s.yRes * s.cropWindow[3]);
    g.sampleBounds[0] = (int)std::floor((float)g.cropped[0] + 0.5f - s.filterRadius[0]);
    g.sampleBounds[1] = (int)std::floor((float)g.cropped[1] + 0.5f - s.filterRadius[1]);
    g.sampleBounds[2] = (int)std::ceil((float)g.cropped[2] - 0.5f + s.filterRadius[0]);
    g.sampleBounds[3] = (int)std::ceil((float)g.cropped[3] - 0.5f + s.filterRadius[1]);
    return g;
}

static float evalFilter(const IileRenderSetup &s, float x, float y) {
    const float rx = s.filterRadius[0], ry = s.filterRadius[1];
    switch (s.filterType) {
    case IILE_FILTER_BOX:
    case IILE_FILTER_MONITOR:
        return 1.f;
    case IILE_FILTER_GAUSSIAN: {
        float alpha = s.filterParam[0];
        float expX = std::exp(-alpha * rx * rx), expY = std::exp(-alpha * ry * ry);
        float gx = std::max(0.f, float(std::exp(-alpha * x * x) - expX));
        float gy = std::max(0.f, float(std::exp(-alpha * y * y) - expY));
        return gx * gy;
    }
    case IILE_FILTER_MITCHELL: {
        float B = s.filterParam[0], C = s.filterParam[1];
        auto m1 = [B, C](float v) {
            v = std::abs(2 * v);
            if (v > 1)
                return ((-B - 6 * C) * v * v * v + (6 * B + 30 * C) * v * v + (-12 * B - 48 * C) * v +
                        (8 * B + 24 * C)) * (1.f / 6.f);
            return ((12 - 9 * B - 6 * C) * v * v * v + (-18 + 12 * B + 6 * C) * v * v + (6 - 2 * B)) *
                   (1.f / 6.f);
        };
        return m1(x * (1 / rx)) * m1(y * (1 / ry));
    }
    case IILE_FILTER_SINC: {
        float tau = s.filterParam[0];
        const float pi = 3.14159265358979323846f;
        auto sinc = [pi](float v) {
            v = std::abs(v);
            if (v < 1e-5) return 1.f;
            return std::sin(pi * v) / (pi * v);
        };
        auto ws = [&](float v, float r) {
            v = std::abs(v);
            if (v > r) return 0.f;
            float lanczos = sinc(v / tau);
            return sinc(v) * lanczos;
        };
        return ws(x, rx) * ws(y, ry);
    }
    case IILE_FILTER_TRIANGLE:
        return std::max(0.f, rx - std::abs(x)) * std::max(0.f, ry - std::abs(y));
    }
    throw std::runtime_error("unknown pixel filter type");
}

void buildFilterTable(const IileRenderSetup &s, float table[kFilterTableWidth * kFilterTableWidth]) {
    int off = 0;
    for (int y = 0; y < kFilterTableWidth; ++y)
        for (int x = 0; x < kFilterTableWidth; ++x, ++off) {
            float px = (x + 0.5f) * s.filterRadius[0] / kFilterTableWidth;
            float py = (y + 0.5f) * s.filterRadius[1] / kFilterTableWidth;
            table[off] = evalFilter(s, px, py);
        }
}

void resolveFilm(const IileRenderSetup &s, const float *film, size_t nPixels, float *rgb) {
    for (size_t i = 0; i < nPixels; ++i) {
        const float r = film[4 * i], g = film[4 * i + 1], b = film[4 * i + 2], w = film[4 * i + 3];
        // RGBToXYZ, spectrum.h:62-66 (the tile merge)
        float X = 0.412453f * r + 0.357580f * g + 0.180423f * b;
        float Y = 0.212671f * r + 0.715160f * g + 0.072169f * b;
        float Z = 0.019334f * r + 0.119193f * g + 0.950227f * b;
        // XYZToRGB, spectrum.h:56-60
        float o[3];
        o[0] = 3.240479f * X - 1.537150f * Y - 0.498535f * Z;
        o[1] = -0.969256f * X + 1.875991f * Y + 0.041556f * Z;
        o[2] = 0.055648f * X - 0.204043f * Y + 1.057311f * Z;
        if (w != 0) {
            float inv = 1.f / w;
            for (int c = 0; c < 3; ++c) o[c] = std::max(0.f, o[c] * inv);
        }
        for (int c = 0; c < 3; ++c) {
            o[c] += 1.f * 0.f;  // splat term: the path integrator never splats
            rgb[3 * i + c] = o[c] * s.filmScale;
        }
    }
}

// ---- image files ------------------------------------------------------------
static bool hasExt(const std::string &p, const char *ext) {
    size_t n = std::strlen(ext);
    if (p.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower((unsigned char)p[p.size() - n + i]) != ext[i]) return false;
    return true;
}

static void writePFM(const std::string &path, const float *rgb, int w, int h) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Unable to open output PFM file \"" + path + "\"");
    // header: "PF", resolution, negative scale = little endian; rows bottom-to-top
    bool ok = std::fprintf(f, "PF\n%d %d\n%f\n", w, h, -1.f) > 0;
    for (int y = h - 1; ok && y >= 0; --y)
        ok = std::fwrite(rgb + (size_t)3 * y * w, sizeof(float), (size_t)3 * w, f) == (size_t)3 * w;
    std::fclose(f);
    if (!ok) throw std::runtime_error("Error writing PFM file \"" + path + "\"");
}

// IEEE binary32 -> binary16, round to nearest even (what OpenEXR's half(float) does)
static uint16_t toHalf(float v) {
    uint32_t x;
    std::memcpy(&x, &v, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? (0x200u | (man >> 13)) : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        int shift = 14 - exp;
        uint32_t h = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;  // may carry into the exponent: correct
    return (uint16_t)(sign | h);
}

static void writeEXR(const std::string &path, const float *rgb, int w, int h, int fullX, int fullY,
                     int offX, int offY) {
    std::vector<uint8_t> hd;
    auto put = [&](const void *p, size_t n) { hd.insert(hd.end(), (const uint8_t *)p, (const uint8_t *)p + n); };
    auto putStr = [&](const char *s) { put(s, std::strlen(s) + 1); };
    auto putI = [&](int32_t v) { put(&v, 4); };
    auto putF = [&](float v) { put(&v, 4); };
    auto attr = [&](const char *name, const char *type, int32_t size) { putStr(name); putStr(type); putI(size); };
    const uint8_t magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    put(magic, 8);
    attr("channels", "chlist", 4 * 18 + 1);
    for (const char *c : {"A", "B", "G", "R"}) {
        putStr(c);
        putI(1);  // HALF
        uint8_t lin[4] = {0, 0, 0, 0};
        put(lin, 4);
        putI(1);
        putI(1);
    }
    hd.push_back(0);
    attr("compression", "compression", 1);
    hd.push_back(0);
    attr("dataWindow", "box2i", 16);
    putI(offX); putI(offY); putI(offX + w - 1); putI(offY + h - 1);
    attr("displayWindow", "box2i", 16);
    putI(0); putI(0); putI(fullX - 1); putI(fullY - 1);
    attr("lineOrder", "lineOrder", 1);
    hd.push_back(0);
    attr("pixelAspectRatio", "float", 4);
    putF(1.f);
    attr("screenWindowCenter", "v2f", 8);
    putF(0.f); putF(0.f);
    attr("screenWindowWidth", "float", 4);
    putF(1.f);
    hd.push_back(0);
    const uint64_t lineBytes = 8 + (uint64_t)w * 4 * 2;
    uint64_t base = hd.size() + (uint64_t)h * 8;
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Error writing \"" + path + "\": cannot open");
    bool ok = std::fwrite(hd.data(), 1, hd.size(), f) == hd.size();
    for (int y = 0; ok && y < h; ++y) {
        uint64_t off = base + (uint64_t)y * lineBytes;
        ok = std::fwrite(&off, 8, 1, f) == 1;
    }
    std::vector<uint16_t> line((size_t)w * 4);
    const uint16_t one = toHalf(1.f);
    for (int y = 0; ok && y < h; ++y) {
        int32_t hdr[2] = {offY + y, (int32_t)(w * 8)};
        for (int x = 0; x < w; ++x) {
            const float *p = rgb + 3 * ((size_t)y * w + x);
            line[x] = one;                    // A
            line[(size_t)w + x] = toHalf(p[2]);       // B
            line[(size_t)2 * w + x] = toHalf(p[1]);   // G
            line[(size_t)3 * w + x] = toHalf(p[0]);   // R
        }
        ok = std::fwrite(hdr, 4, 2, f) == 2 && std::fwrite(line.data(), 2, line.size(), f) == line.size();
    }
    std::fclose(f);
    if (!ok) throw std::runtime_error("Error writing \"" + path + "\"");
}

void writeImage(const std::string &path, const float *rgb, int xRes, int yRes, int fullX, int fullY,
                int offX, int offY) {
    if (hasExt(path, ".pfm")) writePFM(path, rgb, xRes, yRes);
    else if (hasExt(path, ".exr")) writeEXR(path, rgb, xRes, yRes, fullX, fullY, offX, offY);
    else
        throw std::runtime_error("Can't determine image file type from suffix of filename \"" + path +
                                 "\" (the B200 path writes .pfm and .exr)");
}

}  // namespace iile
