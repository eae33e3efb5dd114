// Image textures on the host: ReadImage for PNG / TGA / PFM (src/core/imageio.cpp:60-79, 216-287, 289-400),
// ImageTexture::GetTexture's conversion (src/textures/imagemap.cpp:53-105: flip in y, scale, inverse gamma,
// luminance for float textures) and the MIPMap pyramid (src/core/mipmap.h:100-237: Lanczos resampling to
// powers of two, 2x2 box filter per level with the wrap mode's texel addressing).  The pyramid is built once
// per (file, channels, wrap, scale, gamma) and uploaded as it is; the lookups run on the device
// (cuda/texture.cuh).  PNG decoding is written against the PNG specification on top of zlib's inflate (the
// reference drives the vendored lodepng): 8-bit non-interlaced images of every colour type, reduced to RGB
// as lodepng_decode24 does.
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <tuple>

#include "host_scene.h"

namespace iile {

namespace {

struct RGBImage {
    int w = 0, h = 0;
    std::vector<float> rgb;  // row-major from the top row, 3 floats per pixel
};

std::vector<unsigned char> slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Unable to read image \"" + path + "\"");
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    return std::vector<unsigned char>(s.begin(), s.end());
}

bool hasExtension(const std::string &name, const char *ext) {
    const size_t n = std::strlen(ext);
    if (name.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower((unsigned char)name[name.size() - n + i]) != ext[i]) return false;
    return true;
}

uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

RGBImage readPNG(const std::string &path) {
    const std::vector<unsigned char> d = slurp(path);
    static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8)) throw std::runtime_error(path + ": not a PNG file");
    uint32_t w = 0, h = 0;
    int depth = 0, colour = -1, interlace = 0;
    std::vector<unsigned char> idat, palette;
    for (size_t at = 8; at + 12 <= d.size();) {
        const uint32_t len = be32(&d[at]);
        const std::string type(reinterpret_cast<const char *>(&d[at + 4]), 4);
        if (at + 12 + (size_t)len > d.size()) throw std::runtime_error(path + ": truncated PNG chunk");
        const unsigned char *body = &d[at + 8];
        if (type == "IHDR") {
            if (len < 13) throw std::runtime_error(path + ": bad IHDR");
            w = be32(body), h = be32(body + 4);
            depth = body[8], colour = body[9], interlace = body[12];
        } else if (type == "PLTE")
            palette.assign(body, body + len);
        else if (type == "IDAT")
            idat.insert(idat.end(), body, body + len);
        else if (type == "IEND")
            break;
        at += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || colour < 0) throw std::runtime_error(path + ": PNG without IHDR");
    if (w > 65536 || h > 65536) throw std::runtime_error(path + ": PNG larger than 65536 pixels a side");
    if (depth != 8 || interlace != 0)
        throw std::runtime_error(path + ": only 8-bit non-interlaced PNG images are supported by the B200 path");
    int channels;
    switch (colour) {
    case 0: channels = 1; break;  // grey
    case 2: channels = 3; break;  // RGB
    case 3: channels = 1; break;  // palette index
    case 4: channels = 2; break;  // grey + alpha
    case 6: channels = 4; break;  // RGBA
    default: throw std::runtime_error(path + ": unknown PNG colour type");
    }
    const size_t stride = (size_t)w * channels;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size())
        throw std::runtime_error(path + ": PNG image data does not inflate to the size IHDR announces");
    // undo the scanline filters (PNG specification, section 9)
    std::vector<unsigned char> img(stride * h);
    for (uint32_t y = 0; y < h; ++y) {
        const unsigned char *src = &raw[(stride + 1) * y + 1];
        unsigned char *dst = &img[stride * y];
        const unsigned char *up = y ? &img[stride * (y - 1)] : nullptr;
        const int filter = raw[(stride + 1) * y];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)channels ? dst[i - channels] : 0;
            const int b = up ? up[i] : 0;
            const int c = (up && i >= (size_t)channels) ? up[i - channels] : 0;
            int pred = 0;
            switch (filter) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) >> 1; break;
            case 4: {
                const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                break;
            }
            default: throw std::runtime_error(path + ": unknown PNG filter type");
            }
            dst[i] = (unsigned char)(src[i] + pred);
        }
    }
    RGBImage out;
    out.w = (int)w;
    out.h = (int)h;
    out.rgb.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        unsigned char r, g, b;
        const unsigned char *px = &img[i * channels];
        if (colour == 3) {
            if ((size_t)px[0] * 3 + 2 >= palette.size()) throw std::runtime_error(path + ": PNG palette index out of range");
            r = palette[3 * px[0]], g = palette[3 * px[0] + 1], b = palette[3 * px[0] + 2];
        } else if (channels <= 2)
            r = g = b = px[0];
        else
            r = px[0], g = px[1], b = px[2];
        out.rgb[3 * i] = r / 255.f;  // imageio.cpp:275-278
        out.rgb[3 * i + 1] = g / 255.f;
        out.rgb[3 * i + 2] = b / 255.f;
    }
    return out;
}

// imageio.cpp:216-256 (through the vendored targa.c): true-colour and grey images, raw or run-length encoded
RGBImage readTGA(const std::string &path) {
    const std::vector<unsigned char> d = slurp(path);
    if (d.size() < 18) throw std::runtime_error(path + ": not a TGA file");
    const int idLen = d[0], mapType = d[1], type = d[2];
    const int w = d[12] | (d[13] << 8), h = d[14] | (d[15] << 8), bpp = d[16], desc = d[17];
    if (mapType != 0 || !(type == 2 || type == 3 || type == 10 || type == 11))
        throw std::runtime_error(path + ": colour-mapped TGA images are not supported by the B200 path");
    const int bytes = bpp / 8;
    if (!((type == 2 || type == 10) ? (bytes == 3 || bytes == 4) : bytes == 1))
        throw std::runtime_error(path + ": unsupported TGA pixel depth");
    if (w == 0 || h == 0) throw std::runtime_error(path + ": TGA image without pixels");  // found by tools/fuzz_parser.py
    std::vector<unsigned char> px((size_t)w * h * bytes);
    size_t at = 18 + (size_t)idLen;
    if (type == 2 || type == 3) {
        if (at + px.size() > d.size()) throw std::runtime_error(path + ": truncated TGA file");
        std::memcpy(px.data(), &d[at], px.size());
    } else {
        size_t out = 0;
        while (out < px.size()) {
            if (at >= d.size()) throw std::runtime_error(path + ": truncated TGA file");
            const int head = d[at++], count = (head & 127) + 1;
            if (head & 128) {
                if (at + bytes > d.size()) throw std::runtime_error(path + ": truncated TGA file");
                for (int k = 0; k < count && out < px.size(); ++k, out += bytes) std::memcpy(&px[out], &d[at], bytes);
                at += bytes;
            } else {
                const size_t n = (size_t)count * bytes;
                if (at + n > d.size() || out + n > px.size()) throw std::runtime_error(path + ": truncated TGA file");
                std::memcpy(&px[out], &d[at], n);
                at += n;
                out += n;
            }
        }
    }
    const bool rightToLeft = desc & 0x10, topToBottom = desc & 0x20;
    RGBImage img;
    img.w = w;
    img.h = h;
    img.rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int sx = rightToLeft ? w - 1 - x : x, sy = topToBottom ? y : h - 1 - y;
            const unsigned char *p = &px[((size_t)sy * w + sx) * bytes];
            float *o = &img.rgb[((size_t)y * w + x) * 3];
            if (bytes == 1) o[0] = o[1] = o[2] = p[0] / 255.f;
            else {  // stored B, G, R
                o[0] = p[2] / 255.f;
                o[1] = p[1] / 255.f;
                o[2] = p[0] / 255.f;
            }
        }
    return img;
}

// imageio.cpp:289-400: "PF" (colour) or "Pf" (grey), width height, scale (negative: little endian), rows
// bottom to top
RGBImage readPFM(const std::string &path) {
    const std::vector<unsigned char> d = slurp(path);
    size_t at = 0;
    auto word = [&]() {
        while (at < d.size() && std::isspace(d[at])) ++at;
        std::string w;
        while (at < d.size() && !std::isspace(d[at])) w.push_back((char)d[at++]);
        return w;
    };
    const std::string magic = word();
    if (magic != "PF" && magic != "Pf") throw std::runtime_error(path + ": not a PFM file");
    const int channels = magic == "PF" ? 3 : 1;
    const int w = std::atoi(word().c_str()), h = std::atoi(word().c_str());
    const float scale = (float)std::atof(word().c_str());
    ++at;  // the single whitespace byte after the header
    if (w <= 0 || h <= 0 || w > 65536 || h > 65536 || at + (size_t)w * h * channels * 4 > d.size())
        throw std::runtime_error(path + ": truncated PFM file");
    const bool little = scale < 0;
    RGBImage img;
    img.w = w;
    img.h = h;
    img.rgb.resize((size_t)w * h * 3);
    const float mag = std::fabs(scale);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) {
                const unsigned char *p = &d[at + (((size_t)(h - 1 - y) * w + x) * channels + (channels == 3 ? c : 0)) * 4];
                unsigned char b[4] = {p[0], p[1], p[2], p[3]};
                if (!little) std::swap(b[0], b[3]), std::swap(b[1], b[2]);
                float v;
                std::memcpy(&v, b, 4);
                img.rgb[((size_t)y * w + x) * 3 + c] = mag != 1.f ? v * mag : v;
            }
    return img;
}

float inverseGammaCorrect(float value) {  // pbrt.h:295-298
    if (value <= 0.04045f) return value * 1.f / 12.92f;
    return std::pow((value + 0.055f) * 1.f / 1.055f, 2.4f);
}

float lanczos(float x, float tau = 2) {  // texture.cpp:254-262
    x = std::abs(x);
    if (x < 1e-5f) return 1;
    if (x > 1.f) return 0;
    x *= 3.14159265358979323846f;
    const float s = std::sin(x * tau) / (x * tau);
    const float l = std::sin(x) / x;
    return s * l;
}

int roundUpPow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
int mod(int a, int b) {  // pbrt.h Mod
    const int r = a - (a / b) * b;
    return r < 0 ? r + b : r;
}
int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct ResampleWeight {
    int firstTexel;
    float weight[4];
};
std::vector<ResampleWeight> resampleWeights(int oldRes, int newRes) {  // mipmap.h:78-97
    std::vector<ResampleWeight> wt(newRes);
    const float filterwidth = 2.f;
    for (int i = 0; i < newRes; ++i) {
        const float center = (i + .5f) * oldRes / newRes;
        wt[i].firstTexel = (int)std::floor((center - filterwidth) + 0.5f);
        for (int j = 0; j < 4; ++j) {
            const float pos = wt[i].firstTexel + j + .5f;
            wt[i].weight[j] = lanczos((pos - center) / filterwidth);
        }
        const float invSumWts = 1 / (wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3]);
        for (int j = 0; j < 4; ++j) wt[i].weight[j] *= invSumWts;
    }
    return wt;
}

}  // namespace

int loadImageMap(HostScene *scene, const std::string &file, int channels, int wrap, float scale, bool gamma) {
    HostScene &s = *scene;
    auto &cache = s.imageCache;  // TexInfo cache, imagemap.h:52-75
    const auto key = std::make_tuple(file, channels, wrap, scale, gamma);
    auto hit = cache.find(key);
    if (hit != cache.end()) return hit->second;

    RGBImage img;
    if (hasExtension(file, ".png")) img = readPNG(file);
    else if (hasExtension(file, ".tga")) img = readTGA(file);
    else if (hasExtension(file, ".pfm")) img = readPFM(file);
    else
        throw std::runtime_error("Unable to load image \"" + file + "\": the B200 path reads .png, .tga and .pfm textures");
    // flip in y (imagemap.cpp:73-80), convert (imagemap.h:98-106)
    int resX = img.w, resY = img.h;
    const int C = channels;
    std::vector<float> texels((size_t)resX * resY * C);
    for (int y = 0; y < resY; ++y)
        for (int x = 0; x < resX; ++x) {
            const float *src = &img.rgb[((size_t)(resY - 1 - y) * resX + x) * 3];
            float *dst = &texels[((size_t)y * resX + x) * C];
            if (C == 3)
                for (int c = 0; c < 3; ++c) dst[c] = scale * (gamma ? inverseGammaCorrect(src[c]) : src[c]);
            else {
                const float yv = 0.212671f * src[0] + 0.715160f * src[1] + 0.072169f * src[2];  // RGBSpectrum::y()
                dst[0] = scale * (gamma ? inverseGammaCorrect(yv) : yv);
            }
        }
    // resample to powers of two, mipmap.h:112-171
    if ((resX & (resX - 1)) || (resY & (resY - 1))) {
        const int pw = roundUpPow2(resX), ph = roundUpPow2(resY);
        std::vector<float> res((size_t)pw * ph * C, 0.f);
        const std::vector<ResampleWeight> sW = resampleWeights(resX, pw);
        for (int t = 0; t < resY; ++t)
            for (int sx = 0; sx < pw; ++sx)
                for (int c = 0; c < C; ++c) {
                    float acc = 0.f;
                    for (int j = 0; j < 4; ++j) {
                        int origS = sW[sx].firstTexel + j;
                        if (wrap == IILE_WRAP_REPEAT) origS = mod(origS, resX);
                        else if (wrap == IILE_WRAP_CLAMP) origS = clampi(origS, 0, resX - 1);
                        if (origS >= 0 && origS < resX) acc += sW[sx].weight[j] * texels[((size_t)t * resX + origS) * C + c];
                    }
                    res[((size_t)t * pw + sx) * C + c] = acc;
                }
        const std::vector<ResampleWeight> tW = resampleWeights(resY, ph);
        std::vector<float> work((size_t)ph);
        for (int sx = 0; sx < pw; ++sx)
            for (int c = 0; c < C; ++c) {
                for (int t = 0; t < ph; ++t) {
                    float acc = 0.f;
                    for (int j = 0; j < 4; ++j) {
                        int offset = tW[t].firstTexel + j;
                        if (wrap == IILE_WRAP_REPEAT) offset = mod(offset, resY);
                        else if (wrap == IILE_WRAP_CLAMP) offset = clampi(offset, 0, resY - 1);
                        if (offset >= 0 && offset < resY) acc += tW[t].weight[j] * res[((size_t)offset * pw + sx) * C + c];
                    }
                    work[t] = acc;
                }
                for (int t = 0; t < ph; ++t) res[((size_t)t * pw + sx) * C + c] = work[t] < 0.f ? 0.f : work[t];  // clamp(0, inf)
            }
        texels.swap(res);
        resX = pw;
        resY = ph;
    }
    // the pyramid, mipmap.h:173-195
    int nLevels = 1;
    for (int m = std::max(resX, resY); m > 1; m >>= 1) ++nLevels;
    IileImage im{};
    im.width = resX;
    im.height = resY;
    im.nLevels = nLevels;
    im.channels = C;
    im.texelOffset = s.texels.size();
    s.texels.insert(s.texels.end(), texels.begin(), texels.end());
    int pw = resX, ph = resY;
    size_t prevAt = (size_t)im.texelOffset;
    for (int l = 1; l < nLevels; ++l) {
        const int sRes = std::max(1, pw / 2), tRes = std::max(1, ph / 2);
        const size_t at = s.texels.size();
        s.texels.resize(at + (size_t)sRes * tRes * C);
        auto texel = [&](int sx, int t, int c) -> float {  // MIPMap::Texel on level l - 1, mipmap.h:206-228
            if (wrap == IILE_WRAP_REPEAT) {
                sx = mod(sx, pw);
                t = mod(t, ph);
            } else if (wrap == IILE_WRAP_CLAMP) {
                sx = clampi(sx, 0, pw - 1);
                t = clampi(t, 0, ph - 1);
            } else if (sx < 0 || sx >= pw || t < 0 || t >= ph)
                return 0.f;
            return s.texels[prevAt + ((size_t)t * pw + sx) * C + c];
        };
        for (int t = 0; t < tRes; ++t)
            for (int sx = 0; sx < sRes; ++sx)
                for (int c = 0; c < C; ++c)
                    s.texels[at + ((size_t)t * sRes + sx) * C + c] =
                        .25f * (texel(2 * sx, 2 * t, c) + texel(2 * sx + 1, 2 * t, c) + texel(2 * sx, 2 * t + 1, c) +
                                texel(2 * sx + 1, 2 * t + 1, c));
        prevAt = at;
        pw = sRes;
        ph = tRes;
    }
    s.images.push_back(im);
    return cache[key] = (int)s.images.size() - 1;
}

}  // namespace iile
