// Shape "plymesh": minimal PLY reader (ascii / binary_little_endian /
// binary_big_endian; vertex x y z [nx ny nz] [u v | s t]; face lists of 3 or
// 4 indices, quads split (0,1,2)(3,0,2) as plymesh.cpp:135-146 does) — the subset of
// src/shapes/plymesh.cpp:44-290 (which drives the vendored rply) that
// triangle meshes written by pbrt's own WritePlyFile and by common exporters
// use.  Written against the PLY format description, not rply.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "host_scene.h"

namespace iile {

namespace {
enum class Ty { I8, U8, I16, U16, I32, U32, F32, F64 };
bool parseTy(const std::string &s, Ty *t) {
    if (s == "char" || s == "int8") *t = Ty::I8;
    else if (s == "uchar" || s == "uint8") *t = Ty::U8;
    else if (s == "short" || s == "int16") *t = Ty::I16;
    else if (s == "ushort" || s == "uint16") *t = Ty::U16;
    else if (s == "int" || s == "int32") *t = Ty::I32;
    else if (s == "uint" || s == "uint32") *t = Ty::U32;
    else if (s == "float" || s == "float32") *t = Ty::F32;
    else if (s == "double" || s == "float64") *t = Ty::F64;
    else return false;
    return true;
}
int tySize(Ty t) {
    switch (t) {
    case Ty::I8: case Ty::U8: return 1;
    case Ty::I16: case Ty::U16: return 2;
    case Ty::F64: return 8;
    default: return 4;
    }
}
struct Prop {
    std::string name;
    bool list = false;
    Ty countTy = Ty::U8, ty = Ty::F32;
};
struct Elem {
    std::string name;
    size_t count = 0;
    std::vector<Prop> props;
};
struct Reader {
    std::istream &in;
    int fmt;  // 0 ascii, 1 little, 2 big
    // binary payloads are read into memory once and walked by pointer (an 8M-triangle mesh is
    // 44M scalar reads; through istream::read they cost more than the BVH build)
    std::vector<char> blob;
    size_t at = 0;
    void slurp() {
        if (fmt == 0) return;
        const std::streampos here = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streampos end = in.tellg();
        in.seekg(here);
        blob.resize((size_t)(end - here));
        in.read(blob.data(), (std::streamsize)blob.size());
        if (!in) throw std::runtime_error("plymesh: cannot read the binary payload");
    }
    double get(Ty t) {
        if (fmt == 0) {
            double v;
            if (!(in >> v)) throw std::runtime_error("plymesh: premature end of ascii data");
            return v;
        }
        unsigned char b[8];
        int n = tySize(t);
        if (at + (size_t)n > blob.size()) throw std::runtime_error("plymesh: premature end of binary data");
        std::memcpy(b, blob.data() + at, n);
        at += n;
        if (fmt == 2) std::reverse(b, b + n);
        switch (t) {
        case Ty::I8: return (int8_t)b[0];
        case Ty::U8: return b[0];
        case Ty::I16: { int16_t v; std::memcpy(&v, b, 2); return v; }
        case Ty::U16: { uint16_t v; std::memcpy(&v, b, 2); return v; }
        case Ty::I32: { int32_t v; std::memcpy(&v, b, 4); return v; }
        case Ty::U32: { uint32_t v; std::memcpy(&v, b, 4); return v; }
        case Ty::F32: { float v; std::memcpy(&v, b, 4); return v; }
        default: { double v; std::memcpy(&v, b, 8); return v; }
        }
    }
};
}  // namespace

void loadPly(const std::string &path, std::vector<V3> *P, std::vector<V3> *N, std::vector<float> *UV,
             std::vector<int> *idx) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("Couldn't open PLY file \"" + path + "\"");
    std::string line;
    std::getline(in, line);
    if (line.substr(0, 3) != "ply") throw std::runtime_error(path + ": not a PLY file");
    int fmt = -1;
    std::vector<Elem> elems;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string kw;
        ss >> kw;
        if (kw == "format") {
            std::string f;
            ss >> f;
            fmt = f == "ascii" ? 0 : (f == "binary_little_endian" ? 1 : (f == "binary_big_endian" ? 2 : -1));
        } else if (kw == "element") {
            Elem e;
            ss >> e.name >> e.count;
            elems.push_back(e);
        } else if (kw == "property") {
            if (elems.empty()) throw std::runtime_error(path + ": property before element");
            Prop p;
            std::string t;
            ss >> t;
            if (t == "list") {
                std::string ct, it;
                ss >> ct >> it >> p.name;
                p.list = true;
                if (!parseTy(ct, &p.countTy) || !parseTy(it, &p.ty)) throw std::runtime_error(path + ": bad list type");
            } else {
                if (!parseTy(t, &p.ty)) throw std::runtime_error(path + ": bad property type " + t);
                ss >> p.name;
            }
            elems.back().props.push_back(p);
        } else if (kw == "end_header")
            break;
    }
    if (fmt < 0) throw std::runtime_error(path + ": unknown PLY format");
    Reader rd{in, fmt};
    rd.slurp();
    bool anyN = false, anyUV = false;
    std::vector<int> v;
    for (const Elem &e : elems) {
        if (e.name == "vertex") {
            int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, iu = -1, iv = -1;
            for (size_t k = 0; k < e.props.size(); ++k) {
                const std::string &n = e.props[k].name;
                if (n == "x") ix = k; else if (n == "y") iy = k; else if (n == "z") iz = k;
                else if (n == "nx") inx = k; else if (n == "ny") iny = k; else if (n == "nz") inz = k;
                else if (n == "u" || n == "s" || n == "texture_u" || n == "texture_s") iu = k;
                else if (n == "v" || n == "t" || n == "texture_v" || n == "texture_t") iv = k;
            }
            if (ix < 0 || iy < 0 || iz < 0) throw std::runtime_error(path + ": vertex coordinate property \"x\", \"y\" or \"z\" not found");
            anyN = inx >= 0 && iny >= 0 && inz >= 0;
            anyUV = iu >= 0 && iv >= 0;
            std::vector<double> row(e.props.size());
            P->reserve(P->size() + e.count);
            if (anyN) N->reserve(N->size() + e.count);
            if (anyUV) UV->reserve(UV->size() + 2 * e.count);
            // little-endian rows of float32 scalars (what WritePlyFile and most exporters write): fixed
            // stride, values copied straight out of the payload
            bool allF32 = fmt == 1;
            for (const Prop &p : e.props) allF32 = allF32 && !p.list && p.ty == Ty::F32;
            if (allF32) {
                const size_t stride = 4 * e.props.size();
                if (rd.at + stride * e.count > rd.blob.size()) throw std::runtime_error("plymesh: premature end of binary data");
                const char *src = rd.blob.data() + rd.at;
                auto f32 = [](const char *q) { float v; std::memcpy(&v, q, 4); return v; };
                P->resize(P->size() + e.count);
                V3 *dst = P->data() + (P->size() - e.count);
                for (size_t i = 0; i < e.count; ++i) {
                    const char *r = src + i * stride;
                    dst[i] = V3(f32(r + 4 * ix), f32(r + 4 * iy), f32(r + 4 * iz));
                    if (anyN) N->push_back(V3(f32(r + 4 * inx), f32(r + 4 * iny), f32(r + 4 * inz)));
                    if (anyUV) { UV->push_back(f32(r + 4 * iu)); UV->push_back(f32(r + 4 * iv)); }
                }
                rd.at += stride * e.count;
                continue;
            }
            for (size_t i = 0; i < e.count; ++i) {
                for (size_t k = 0; k < e.props.size(); ++k) {
                    if (e.props[k].list) {
                        int n = (int)rd.get(e.props[k].countTy);
                        for (int j = 0; j < n; ++j) rd.get(e.props[k].ty);
                    } else
                        row[k] = rd.get(e.props[k].ty);
                }
                P->push_back(V3((float)row[ix], (float)row[iy], (float)row[iz]));
                if (anyN) N->push_back(V3((float)row[inx], (float)row[iny], (float)row[inz]));
                if (anyUV) { UV->push_back((float)row[iu]); UV->push_back((float)row[iv]); }
            }
        } else if (e.name == "face") {
            idx->reserve(idx->size() + 3 * e.count);
            // little-endian faces that are one list of 32-bit indices counted by a byte
            if (fmt == 1 && e.props.size() == 1 && e.props[0].list && e.props[0].countTy == Ty::U8 &&
                (e.props[0].ty == Ty::I32 || e.props[0].ty == Ty::U32) &&
                (e.props[0].name == "vertex_indices" || e.props[0].name == "vertex_index")) {
                const unsigned char *q = reinterpret_cast<const unsigned char *>(rd.blob.data()) + rd.at;
                const unsigned char *end = reinterpret_cast<const unsigned char *>(rd.blob.data()) + rd.blob.size();
                // first pass: how many triangles (a quad gives two), so that the output is sized once
                size_t nOut = 0;
                {
                    const unsigned char *w = q;
                    for (size_t i = 0; i < e.count; ++i) {
                        if (w >= end || w + 1 + 4 * (size_t)w[0] > end) throw std::runtime_error("plymesh: premature end of binary data");
                        nOut += w[0] == 3 ? 3 : (w[0] == 4 ? 6 : 0);
                        w += 1 + 4 * (size_t)w[0];
                    }
                }
                const size_t before = idx->size();
                idx->resize(before + nOut);
                int *d = idx->data() + before;
                for (size_t i = 0; i < e.count; ++i) {
                    const unsigned n = q[0];
                    if (n == 3) {
                        std::memcpy(d, q + 1, 12);
                        d += 3;
                    } else if (n == 4) {
                        int v4[4];
                        std::memcpy(v4, q + 1, 16);
                        d[0] = v4[0]; d[1] = v4[1]; d[2] = v4[2]; d[3] = v4[3]; d[4] = v4[0]; d[5] = v4[2];
                        d += 6;
                    }
                    q += 1 + 4 * (size_t)n;
                }
                rd.at = (size_t)(q - reinterpret_cast<const unsigned char *>(rd.blob.data()));
                continue;
            }
            for (size_t i = 0; i < e.count; ++i)
                for (const Prop &p : e.props) {
                    if (!p.list) { rd.get(p.ty); continue; }
                    int n = (int)rd.get(p.countTy);
                    if (n < 0) throw std::runtime_error(path + ": negative list length");
                    v.resize(n);
                    for (int j = 0; j < n; ++j) v[j] = (int)rd.get(p.ty);
                    if (p.name != "vertex_indices" && p.name != "vertex_index") continue;
                    if (n == 3) idx->insert(idx->end(), {v[0], v[1], v[2]});
                    else if (n == 4) idx->insert(idx->end(), {v[0], v[1], v[2], v[3], v[0], v[2]});
                    // other polygon sizes are ignored, as the reference does (plymesh.cpp:112-117)
                }
        } else {
            for (size_t i = 0; i < e.count; ++i)
                for (const Prop &p : e.props) {
                    if (p.list) {
                        int n = (int)rd.get(p.countTy);
                        for (int j = 0; j < n; ++j) rd.get(p.ty);
                    } else
                        rd.get(p.ty);
                }
        }
    }
    if (P->empty() || idx->empty()) throw std::runtime_error(path + ": PLY file is invalid! No face/vertex elements found!");
    for (int v : *idx)
        if (v < 0 || (size_t)v >= P->size()) throw std::runtime_error(path + ": plymesh vertex index out of bounds");
}

}  // namespace iile
