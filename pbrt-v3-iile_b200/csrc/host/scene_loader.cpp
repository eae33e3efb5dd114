// .pbrt front-end for the hot path: tokenizer, typed parameter lists and the
// directive state machine, producing a flattened HostScene.
//
// Mirrors the behaviour of the reference's hand-written parser
// (src/core/parser.cpp:186-257 tokenizer, :343-453 parameter typing,
// :800-1020 directive loop) and of the api.cpp state machine
// (CTM post-multiplication api.cpp:934-937, Camera api.cpp:1113-1124,
// Attribute/Transform stacks api.cpp:1170-1226, Material/Texture/NamedMaterial
// api.cpp:1228-1330, LightSource/AreaLightSource api.cpp:1332-1369, Shape
// api.cpp:1371-1467 incl. the per-shape material override :1469-1532) for the
// directive subset SURVEY §8b lists.  Anything outside that subset raises —
// the reference would render it differently, so silently ignoring it would
// break the drop-in contract.
#include <chrono>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>

#include "host_scene.h"

namespace iile {

namespace {

[[noreturn]] void fail(const std::string &msg) { throw std::runtime_error(msg); }

// ---------------------------------------------------------------- tokenizer
struct Tokenizer {
    std::string text, file;
    size_t pos = 0;
    int line = 1;
    Tokenizer(const std::string &path) : file(path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) fail(path + ": " + std::strerror(errno));
        std::stringstream ss;
        ss << f.rdbuf();
        text = ss.str();
    }
    [[noreturn]] void err(const std::string &m) const {
        fail(file + ":" + std::to_string(line) + ": " + m);
    }
    // Returns false at EOF.  Quoted strings keep their quotes (so the caller
    // can tell `"foo"` from `foo`), escapes are decoded.
    bool next(std::string *tok) {
        tok->clear();
        while (pos < text.size()) {
            char c = text[pos++];
            if (c == '\n') { ++line; continue; }
            if (c == ' ' || c == '\t' || c == '\r') continue;
            if (c == '#') {
                while (pos < text.size() && text[pos] != '\n') ++pos;
                continue;
            }
            if (c == '"') {
                tok->push_back('"');
                while (true) {
                    if (pos >= text.size()) err("premature EOF");
                    char d = text[pos++];
                    if (d == '"') break;
                    if (d == '\n') err("unterminated string");
                    if (d == '\\') {
                        if (pos >= text.size()) err("premature EOF");
                        char e = text[pos++];
                        switch (e) {
                        case 'b': d = '\b'; break;
                        case 'f': d = '\f'; break;
                        case 'n': d = '\n'; break;
                        case 'r': d = '\r'; break;
                        case 't': d = '\t'; break;
                        case '\\': d = '\\'; break;
                        case '\'': d = '\''; break;
                        case '"': d = '"'; break;
                        default: err(std::string("unexpected escaped character \"") + e + "\"");
                        }
                    }
                    tok->push_back(d);
                }
                tok->push_back('"');
                return true;
            }
            if (c == '[' || c == ']') {
                tok->push_back(c);
                return true;
            }
            tok->push_back(c);
            while (pos < text.size()) {
                char d = text[pos];
                if (d == ' ' || d == '\n' || d == '\t' || d == '\r' || d == '"' || d == '[' ||
                    d == ']')
                    break;
                tok->push_back(d);
                ++pos;
            }
            return true;
        }
        return false;
    }
};

bool isQuoted(const std::string &s) { return s.size() >= 2 && s.front() == '"' && s.back() == '"'; }

// parser.cpp:259-303: integers through strtol, everything else through strtof
double parseNumber(const Tokenizer &t, const std::string &s) {
    if (s.size() == 1) {
        if (!(s[0] >= '0' && s[0] <= '9')) t.err("\"" + s + "\": expected a number");
        return s[0] - '0';
    }
    bool isInt = true;
    for (char c : s)
        if (!(c >= '0' && c <= '9')) isInt = false;
    char *end = nullptr;
    double v = isInt ? (double)strtol(s.c_str(), &end, 10) : (double)strtof(s.c_str(), &end);
    if (v == 0 && end == s.c_str()) t.err(s + ": expected a number");
    return v;
}

// ---------------------------------------------------------------- parameters
enum class PT { Int, Bool, Float, Point2, Vector2, Point3, Vector3, Normal, RGB, XYZ, String, Texture };

struct Param {
    PT type;
    std::string name;
    std::vector<double> nums;
    std::vector<std::string> strs;
    mutable bool used = false;
};

struct ParamSet {
    std::vector<Param> items;

    const Param *find(const std::string &name, std::initializer_list<PT> types) const {
        for (const Param &p : items)
            if (p.name == name)
                for (PT t : types)
                    if (p.type == t) {
                        p.used = true;
                        return &p;
                    }
        return nullptr;
    }
    float oneFloat(const std::string &n, float d) const {
        const Param *p = find(n, {PT::Float});
        return (p && p->nums.size() == 1) ? (float)p->nums[0] : d;
    }
    int oneInt(const std::string &n, int d) const {
        const Param *p = find(n, {PT::Int});
        return (p && p->nums.size() == 1) ? (int)p->nums[0] : d;
    }
    bool oneBool(const std::string &n, bool d) const {
        const Param *p = find(n, {PT::Bool});
        return (p && p->strs.size() == 1) ? (p->strs[0] == "true") : d;
    }
    std::string oneString(const std::string &n, const std::string &d) const {
        const Param *p = find(n, {PT::String});
        return (p && p->strs.size() == 1) ? p->strs[0] : d;
    }
    std::string textureName(const std::string &n) const {
        const Param *p = find(n, {PT::Texture});
        return (p && p->strs.size() == 1) ? p->strs[0] : "";
    }
    bool onePoint3(const std::string &n, V3 *out) const {
        const Param *p = find(n, {PT::Point3});
        if (!p || p->nums.size() != 3) return false;
        *out = V3((float)p->nums[0], (float)p->nums[1], (float)p->nums[2]);
        return true;
    }
    // RGBSpectrum::FromRGB is the identity; xyz goes through XYZToRGB (spectrum.h:56-60)
    bool oneSpectrum(const std::string &n, float rgb[3]) const {
        const Param *p = find(n, {PT::RGB, PT::XYZ});
        if (!p || p->nums.size() != 3) return false;
        float a = (float)p->nums[0], b = (float)p->nums[1], c = (float)p->nums[2];
        if (p->type == PT::RGB) {
            rgb[0] = a; rgb[1] = b; rgb[2] = c;
        } else {
            rgb[0] = 3.240479f * a - 1.537150f * b - 0.498535f * c;
            rgb[1] = -0.969256f * a + 1.875991f * b + 0.041556f * c;
            rgb[2] = 0.055648f * a - 0.204043f * b + 1.057311f * c;
        }
        return true;
    }
    const Param *floats(const std::string &n) const { return find(n, {PT::Float}); }
    const Param *ints(const std::string &n) const { return find(n, {PT::Int}); }
};

bool lookupType(const std::string &decl, PT *type, std::string *name, std::string *why) {
    std::istringstream ss(decl);
    std::string t;
    if (!(ss >> t)) { *why = "Parameter \"" + decl + "\" doesn't have a type declaration?!"; return false; }
    static const std::map<std::string, PT> table = {
        {"float", PT::Float},     {"integer", PT::Int},     {"bool", PT::Bool},
        {"point2", PT::Point2},   {"vector2", PT::Vector2}, {"point3", PT::Point3},
        {"vector3", PT::Vector3}, {"point", PT::Point3},    {"vector", PT::Vector3},
        {"normal", PT::Normal},   {"string", PT::String},   {"texture", PT::Texture},
        {"color", PT::RGB},       {"rgb", PT::RGB},         {"xyz", PT::XYZ}};
    auto it = table.find(t);
    if (it == table.end()) {
        if (t == "blackbody" || t == "spectrum")
            *why = "parameter type \"" + t + "\" (in \"" + decl + "\") is not supported by the B200 path";
        else
            *why = "Unable to decode type from \"" + decl + "\"";
        return false;
    }
    *type = it->second;
    if (!(ss >> *name)) { *why = "Unable to find parameter name from \"" + decl + "\""; return false; }
    return true;
}

// ---------------------------------------------------------------- graphics state
struct MaterialDef {
    std::string type = "matte";
    ParamSet params;
    int index = -2;  // index into HostScene::materials once instantiated (-1: none)
};

// A named texture as a material parameter sees it: a constant (folded into the material record, as the
// reference's ConstantTexture would evaluate) or an entry of HostScene::textures.
struct TexRef {
    int index = -1;
    float v[3] = {0, 0, 0};
};

struct GState {
    std::shared_ptr<MaterialDef> material = std::make_shared<MaterialDef>();
    std::map<std::string, std::shared_ptr<MaterialDef>> namedMaterials;
    std::map<std::string, TexRef> floatTextures;
    std::map<std::string, TexRef> spectrumTextures;
    std::string areaLight;
    ParamSet areaLightParams;
    bool reverseOrientation = false;
};

struct Loader {
    HostScene *scene;
    std::string searchDir;
    Xform ctm;
    std::map<std::string, Xform> namedCS;
    GState gs;
    std::vector<GState> gsStack;
    std::vector<Xform> xfStack;
    std::vector<char> pushKind;  // 'a' attribute, 't' transform (parser.cpp mismatched-nesting check)
    bool inWorld = false, sawWorldEnd = false;
    // option block
    std::string cameraName = "perspective", filmName = "image", samplerName = "halton",
                filterName = "box", integratorName = "path", accelName = "bvh";
    ParamSet cameraParams, filmParams, samplerParams, filterParams, integratorParams, accelParams;
    Xform cameraToWorld;

    std::string resolve(const std::string &f) const {
        if (!f.empty() && f[0] == '/') return f;
        return searchDir.empty() ? f : searchDir + "/" + f;
    }

    // ---- textures: TextureParams::GetFloatTexture / GetSpectrumTexture (paramset.cpp:560-620) ----------
    TexRef floatRef(const ParamSet &geom, const ParamSet &mat, const std::string &n, float def) {
        std::string tn = geom.textureName(n);
        if (tn.empty()) tn = mat.textureName(n);
        if (!tn.empty()) {
            auto it = gs.floatTextures.find(tn);
            if (it == gs.floatTextures.end()) fail("Couldn't find float texture named \"" + tn + "\" for parameter \"" + n + "\"");
            return it->second;
        }
        TexRef r;
        r.v[0] = r.v[1] = r.v[2] = geom.oneFloat(n, mat.oneFloat(n, def));
        return r;
    }
    TexRef spectrumRef(const ParamSet &geom, const ParamSet &mat, const std::string &n, float def) {
        std::string tn = geom.textureName(n);
        if (tn.empty()) tn = mat.textureName(n);
        if (!tn.empty()) {
            auto it = gs.spectrumTextures.find(tn);
            if (it == gs.spectrumTextures.end()) fail("Couldn't find spectrum texture named \"" + tn + "\" for parameter \"" + n + "\"");
            return it->second;
        }
        TexRef r;
        if (geom.oneSpectrum(n, r.v)) return r;
        if (mat.oneSpectrum(n, r.v)) return r;
        r.v[0] = r.v[1] = r.v[2] = def;
        return r;
    }
    // parameters that have to be constants here (light spectra, texture "value" / "amount", ...)
    float floatTex(const ParamSet &geom, const ParamSet &mat, const std::string &n, float def) {
        const TexRef r = floatRef(geom, mat, n, def);
        if (r.index >= 0) fail("parameter \"" + n + "\": an image texture is not supported here by the B200 path");
        return r.v[0];
    }
    void spectrumTex(const ParamSet &geom, const ParamSet &mat, const std::string &n,
                     float def, float out[3]) {
        const TexRef r = spectrumRef(geom, mat, n, def);
        if (r.index >= 0) fail("parameter \"" + n + "\": an image texture is not supported here by the B200 path");
        for (int i = 0; i < 3; ++i) out[i] = r.v[i];
    }
    // the texture record of a reference (a constant becomes a ConstantTexture entry)
    int textureIndex(const TexRef &r, int channels) {
        if (r.index >= 0) return r.index;
        IileTexture t;
        std::memset(&t, 0, sizeof(t));
        t.type = IILE_TEX_CONSTANT;
        t.channels = channels;
        for (int i = 0; i < 3; ++i) t.value[i] = r.v[i];
        t.tex1 = t.tex2 = t.image = -1;
        scene->textures.push_back(t);
        return (int)scene->textures.size() - 1;
    }

    // MakeMaterial (api.cpp:432-520) for the four in-scope materials.
    int makeMaterial(const std::string &type, const ParamSet &geom, const ParamSet &mat) {
        if (type == "" || type == "none") return -1;
        IileMaterial m;
        std::memset(&m, 0, sizeof(m));
        IileMaterialTextures mt;
        for (int &x : mt.slot) x = -1;
        bool textured = false;
        // a parameter: the constant goes into the record; a texture goes into the material's slot and is
        // evaluated at every hit (cuda/texture.cuh resolveMaterial)
        auto spec = [&](const char *n, float def, float *dst, int slot) {
            const TexRef r = spectrumRef(geom, mat, n, def);
            for (int i = 0; i < 3; ++i) dst[i] = r.index >= 0 ? def : r.v[i];
            if (r.index >= 0) mt.slot[slot] = r.index, textured = true;
        };
        auto flt = [&](const char *n, float def, float *dst, int slot) {
            const TexRef r = floatRef(geom, mat, n, def);
            *dst = r.index >= 0 ? def : r.v[0];
            if (r.index >= 0) mt.slot[slot] = r.index, textured = true;
        };
        auto given = [&](const char *n) {
            return !geom.textureName(n).empty() || !mat.textureName(n).empty() || geom.floats(n) || mat.floats(n);
        };
        auto remap = [&]() { m.remapRoughness = geom.oneBool("remaproughness", mat.oneBool("remaproughness", true)); };
        if (type == "matte") {  // matte.cpp:64-71
            m.type = IILE_MAT_MATTE;
            spec("Kd", 0.5f, m.Kd, IILE_TEXSLOT_KD);
            flt("sigma", 0.f, &m.sigma, IILE_TEXSLOT_SIGMA);
        } else if (type == "plastic") {  // plastic.cpp:72-83
            m.type = IILE_MAT_PLASTIC;
            spec("Kd", 0.25f, m.Kd, IILE_TEXSLOT_KD);
            spec("Ks", 0.25f, m.Ks, IILE_TEXSLOT_KS);
            flt("roughness", 0.1f, &m.roughness, IILE_TEXSLOT_ROUGHNESS);
            remap();
        } else if (type == "mirror") {  // mirror.cpp:58-64
            m.type = IILE_MAT_MIRROR;
            spec("Kr", 0.9f, m.Ks, IILE_TEXSLOT_KS);
        } else if (type == "glass") {  // glass.cpp:94-110
            m.type = IILE_MAT_GLASS;
            spec("Kr", 1.f, m.Ks, IILE_TEXSLOT_KS);
            spec("Kt", 1.f, m.Kt, IILE_TEXSLOT_KT);
            // "eta" wins over "index" when both are given: glass.cpp:99-100
            flt(given("eta") ? "eta" : "index", 1.5f, &m.eta, IILE_TEXSLOT_ETA);
            flt("uroughness", 0.f, &m.uRoughness, IILE_TEXSLOT_UROUGHNESS);
            flt("vroughness", 0.f, &m.vRoughness, IILE_TEXSLOT_VROUGHNESS);
            if (mt.slot[IILE_TEXSLOT_ETA] >= 0 || mt.slot[IILE_TEXSLOT_UROUGHNESS] >= 0 || mt.slot[IILE_TEXSLOT_VROUGHNESS] >= 0)
                fail("glass: image textures for \"index\" / \"uroughness\" / \"vroughness\" are not supported by the B200 path");
            remap();
        } else if (type == "substrate") {  // substrate.cpp:67-83
            m.type = IILE_MAT_SUBSTRATE;
            spec("Kd", 0.5f, m.Kd, IILE_TEXSLOT_KD);
            spec("Ks", 0.5f, m.Ks, IILE_TEXSLOT_KS);
            flt("uroughness", 0.1f, &m.uRoughness, IILE_TEXSLOT_UROUGHNESS);
            flt("vroughness", 0.1f, &m.vRoughness, IILE_TEXSLOT_VROUGHNESS);
            remap();
        } else if (type == "translucent") {  // translucent.cpp:84-105
            m.type = IILE_MAT_TRANSLUCENT;
            spec("Kd", 0.25f, m.Kd, IILE_TEXSLOT_KD);
            spec("Ks", 0.25f, m.Ks, IILE_TEXSLOT_KS);
            spec("reflect", 0.5f, m.Kr, IILE_TEXSLOT_KR);
            spec("transmit", 0.5f, m.Kt, IILE_TEXSLOT_KT);
            flt("roughness", 0.1f, &m.roughness, IILE_TEXSLOT_ROUGHNESS);
            remap();
        } else if (type == "uber") {  // uber.cpp:106-135
            m.type = IILE_MAT_UBER;
            spec("Kd", 0.25f, m.Kd, IILE_TEXSLOT_KD);
            spec("Ks", 0.25f, m.Ks, IILE_TEXSLOT_KS);
            spec("Kr", 0.f, m.Kr, IILE_TEXSLOT_KR);
            spec("Kt", 0.f, m.Kt, IILE_TEXSLOT_KT);
            flt("roughness", 0.1f, &m.roughness, IILE_TEXSLOT_ROUGHNESS);
            // uber.cpp:71-79: u from "uroughness" else "roughness"; v from "vroughness" else u
            if (given("uroughness")) flt("uroughness", 0.f, &m.uRoughness, IILE_TEXSLOT_UROUGHNESS);
            else m.uRoughness = m.roughness, mt.slot[IILE_TEXSLOT_UROUGHNESS] = mt.slot[IILE_TEXSLOT_ROUGHNESS];
            if (given("vroughness")) flt("vroughness", 0.f, &m.vRoughness, IILE_TEXSLOT_VROUGHNESS);
            else m.vRoughness = m.uRoughness, mt.slot[IILE_TEXSLOT_VROUGHNESS] = mt.slot[IILE_TEXSLOT_UROUGHNESS];
            flt(given("eta") ? "eta" : "index", 1.5f, &m.eta, IILE_TEXSLOT_ETA);
            if (mt.slot[IILE_TEXSLOT_ETA] >= 0) fail("uber: an image texture for \"index\" is not supported by the B200 path");
            spec("opacity", 1.f, m.opacity, IILE_TEXSLOT_OPACITY);
            remap();
        } else {
            fail("Material \"" + type + "\" is not supported by the B200 path (matte/plastic/mirror/glass/uber/substrate/translucent only)");
        }
        // "bumpmap" (a float texture, every material: e.g. matte.cpp:67): Material::Bump runs even for a
        // constant displacement (it rebuilds the shading frame from dndu / dndv)
        {
            std::string bn = geom.textureName("bumpmap");
            if (bn.empty()) bn = mat.textureName("bumpmap");
            if (!bn.empty()) {
                auto it = gs.floatTextures.find(bn);
                if (it == gs.floatTextures.end()) fail("Couldn't find float texture named \"" + bn + "\" for parameter \"bumpmap\"");
                mt.slot[IILE_TEXSLOT_BUMP] = textureIndex(it->second, 1);
                textured = true;
            }
        }
        if (textured) {
            scene->materialTextures.push_back(mt);
            m.texSet = (int)scene->materialTextures.size();
        }
        scene->materials.push_back(m);
        return (int)scene->materials.size() - 1;
    }

    int currentMaterialIndex() {
        MaterialDef &d = *gs.material;
        if (d.index == -2) d.index = makeMaterial(d.type, d.params, ParamSet());
        return d.index;
    }

    // shapeMaySetMaterialParameters, api.cpp:1469-1517
    static bool shapeMaySetMaterial(const ParamSet &ps) {
        for (const Param &p : ps.items) {
            switch (p.type) {
            case PT::Texture:
                if (p.name != "alpha" && p.name != "shadowalpha") return true;
                break;
            case PT::Float:
                if (p.nums.size() == 1 && p.name != "radius") return true;
                break;
            case PT::String:
                if (p.strs.size() == 1 && p.name != "filename" && p.name != "type" && p.name != "scheme")
                    return true;
                break;
            case PT::Bool:
                if (p.strs.size() == 1) return true;
                break;
            case PT::Int:
                if (p.nums.size() == 1) return true;
                break;
            case PT::Point2:
            case PT::Vector2:
                if (p.nums.size() == 2) return true;
                break;
            default:  // point3 / vector3 / normal / spectra
                if (p.nums.size() == 3) return true;
                break;
            }
        }
        return false;
    }

    // ---- shapes ---------------------------------------------------------
    void addMesh(const std::vector<V3> &P, const std::vector<V3> *N, const std::vector<float> *UV,
                 const std::vector<int> &idx, int material) {
        HostScene &s = *scene;
        uint32_t base = s.nVerts();
        // TriangleMesh ctor, triangle.cpp:72-88: positions and normals go to world space
        s.P.reserve(s.P.size() + 3 * P.size());
        s.prims.reserve(s.prims.size() + idx.size() / 3);
        for (const V3 &p : P) {
            V3 w = ctm.point(p);
            s.P.push_back(w.x); s.P.push_back(w.y); s.P.push_back(w.z);
        }
        uint32_t flags = 0;
        if (N) {
            if (s.N.empty()) s.N.assign((size_t)base * 3, 0.f);
            for (const V3 &n : *N) {
                V3 w = ctm.normal(n);
                s.N.push_back(w.x); s.N.push_back(w.y); s.N.push_back(w.z);
            }
            flags |= IILE_PRIM_HAS_N;
        }
        if (UV) {
            if (s.UV.empty()) s.UV.assign((size_t)base * 2, 0.f);
            s.UV.insert(s.UV.end(), UV->begin(), UV->end());
            flags |= IILE_PRIM_HAS_UV;
        }
        s.padVertexPools();
        if (gs.reverseOrientation ^ ctm.swapsHandedness()) flags |= IILE_PRIM_FLIP_NORMAL;
        size_t nTris = idx.size() / 3;
        for (size_t t = 0; t < nTris; ++t) {
            IilePrim pr;
            std::memset(&pr, 0, sizeof(pr));
            pr.type = IILE_PRIM_TRIANGLE;
            for (int k = 0; k < 3; ++k) {
                int v = idx[3 * t + k];
                if (v < 0 || v >= (int)P.size()) fail("trianglemesh has out-of-bounds vertex index");
                pr.v[k] = (int)base + v;
            }
            pr.material = material;
            pr.light = -1;
            pr.flags = flags;
            addPrim(pr);
        }
    }

    void addPrim(IilePrim pr) {
        HostScene &s = *scene;
        int primIndex = (int)s.prims.size();
        if (!gs.areaLight.empty()) {
            // MakeAreaLight / CreateDiffuseAreaLight, api.cpp:764-774, diffuse.cpp:135-146:
            // one light per shape, i.e. per triangle.
            if (gs.areaLight != "area" && gs.areaLight != "diffuse")
                fail("Area light \"" + gs.areaLight + "\" unknown.");
            IileLight L;
            std::memset(&L, 0, sizeof(L));
            L.type = IILE_LIGHT_DIFFUSE_AREA;
            L.prim = primIndex;
            float Le[3] = {1, 1, 1}, sc[3] = {1, 1, 1};
            gs.areaLightParams.oneSpectrum("L", Le);
            gs.areaLightParams.oneSpectrum("scale", sc);
            for (int i = 0; i < 3; ++i) L.L[i] = Le[i] * sc[i];
            L.twoSided = gs.areaLightParams.oneBool("twosided", false) ? 1 : 0;
            // diffuse.cpp:140-141: "samples", else "nsamples", else 1
            const int nSamples = gs.areaLightParams.oneInt("samples", gs.areaLightParams.oneInt("nsamples", 1));
            if (nSamples < 1 || nSamples > 65536) fail("area light \"samples\" out of range");
            L.twoSided |= (nSamples - 1) << 8;
            pr.light = (int)s.lights.size();
            s.prims.push_back(pr);
            L.area = primArea(s, pr);
            s.lights.push_back(L);
        } else {
            s.prims.push_back(pr);
        }
    }

    void shape(const std::string &name, const ParamSet &ps) {
        if (!inWorld) fail("Scene description must be inside world block; \"Shape\" not allowed.");
        // GetMaterialForShape, api.cpp:1519-1532
        int material;
        if (shapeMaySetMaterial(ps))
            material = makeMaterial(gs.material->type, ps, gs.material->params);
        else
            material = currentMaterialIndex();
        if (!ps.textureName("alpha").empty() || !ps.textureName("shadowalpha").empty() ||
            ps.floats("alpha") || ps.floats("shadowalpha"))
            fail("alpha / shadowalpha masks are not supported by the B200 path");

        if (name == "trianglemesh") {  // CreateTriangleMeshShape, triangle.cpp:619-715
            const Param *pi = ps.ints("indices");
            const Param *pp = ps.find("P", {PT::Point3});
            if (!pp) fail("Vertex positions \"P\" not provided with triangle mesh shape");
            std::vector<int> idx;
            if (pi) for (double v : pi->nums) idx.push_back((int)v);
            else if (pp->nums.size() == 9) idx = {0, 1, 2};
            else fail("Vertex indices \"indices\" not provided with triangle mesh shape");
            if (idx.size() % 3) idx.resize(idx.size() - idx.size() % 3);
            std::vector<V3> P;
            for (size_t i = 0; i + 2 < pp->nums.size(); i += 3)
                P.push_back(V3((float)pp->nums[i], (float)pp->nums[i + 1], (float)pp->nums[i + 2]));
            std::vector<float> uv;
            bool haveUV = false;
            const Param *puv = ps.find("uv", {PT::Point2});
            if (!puv) puv = ps.find("st", {PT::Point2});
            if (!puv) puv = ps.floats("uv");
            if (!puv) puv = ps.floats("st");
            if (puv) {
                if (puv->nums.size() == 2 * P.size()) {
                    for (double v : puv->nums) uv.push_back((float)v);
                    haveUV = true;
                } else if (puv->nums.size() > 2 * P.size()) {
                    for (size_t i = 0; i < 2 * P.size(); ++i) uv.push_back((float)puv->nums[i]);
                    haveUV = true;
                }  // fewer: discarded with a warning in the reference
            }
            if (ps.find("S", {PT::Vector3})) fail("trianglemesh tangents \"S\" are not supported by the B200 path");
            std::vector<V3> N;
            bool haveN = false;
            const Param *pn = ps.find("N", {PT::Normal});
            if (pn && pn->nums.size() == 3 * P.size()) {
                for (size_t i = 0; i + 2 < pn->nums.size(); i += 3)
                    N.push_back(V3((float)pn->nums[i], (float)pn->nums[i + 1], (float)pn->nums[i + 2]));
                haveN = true;
            }
            addMesh(P, haveN ? &N : nullptr, haveUV ? &uv : nullptr, idx, material);
        } else if (name == "loopsubdiv") {  // CreateLoopSubdiv, loopsubdiv.cpp:402-424
            int nLevels = ps.oneInt("levels", ps.oneInt("nlevels", 3));
            const Param *pi = ps.ints("indices");
            const Param *pp = ps.find("P", {PT::Point3});
            if (!pi) fail("Vertex indices \"indices\" not provided for LoopSubdiv shape.");
            if (!pp) fail("Vertex positions \"P\" not provided for LoopSubdiv shape.");
            std::vector<int> idx;
            for (double v : pi->nums) idx.push_back((int)v);
            std::vector<V3> P;
            for (size_t i = 0; i + 2 < pp->nums.size(); i += 3)
                P.push_back(V3((float)pp->nums[i], (float)pp->nums[i + 1], (float)pp->nums[i + 2]));
            std::vector<V3> oP, oN;
            std::vector<int> oI;
            loopSubdivide(nLevels, idx, P, &oP, &oN, &oI);
            addMesh(oP, &oN, nullptr, oI, material);
        } else if (name == "plymesh") {
            std::string fn = ps.oneString("filename", "");
            if (fn.empty()) fail("plymesh: no \"filename\"");
            std::vector<V3> P, N;
            std::vector<float> UV;
            std::vector<int> idx;
            const auto tp0 = std::chrono::steady_clock::now();
            loadPly(resolve(fn), &P, &N, &UV, &idx);
            const auto tp1 = std::chrono::steady_clock::now();
            addMesh(P, N.empty() ? nullptr : &N, UV.empty() ? nullptr : &UV, idx, material);
            if (std::getenv("IILE_BUILD_TRACE"))
                std::fprintf(stderr, "plymesh: read %.3f s, flatten %.3f s\n", std::chrono::duration<double>(tp1 - tp0).count(),
                             std::chrono::duration<double>(std::chrono::steady_clock::now() - tp1).count());
        } else if (name == "cylinder") {  // CreateCylinderShape, cylinder.cpp:222-233; ctor cylinder.h:49-57
            IileSphere sp;
            std::memset(&sp, 0, sizeof(sp));
            std::memcpy(sp.objectToWorld, ctm.m.m, sizeof(float) * 16);
            std::memcpy(sp.worldToObject, ctm.inv.m, sizeof(float) * 16);
            auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
            sp.radius = ps.oneFloat("radius", 1.f);
            const float zmin = ps.oneFloat("zmin", -1.f), zmax = ps.oneFloat("zmax", 1.f);
            sp.zMin = std::min(zmin, zmax);
            sp.zMax = std::max(zmin, zmax);
            const float phimax = ps.oneFloat("phimax", 360.f);
            sp.phiMax = radians(clampf(phimax, 0, 360));
            sp.phiMaxDegrees = phimax;
            sp.flags = (gs.reverseOrientation ? 1u : 0u) | (ctm.swapsHandedness() ? 2u : 0u) | 8u;
            scene->spheres.push_back(sp);
            IilePrim pr;
            std::memset(&pr, 0, sizeof(pr));
            pr.type = IILE_PRIM_SPHERE;
            pr.v[0] = (int)scene->spheres.size() - 1;
            pr.material = material;
            pr.light = -1;
            addPrim(pr);
        } else if (name == "disk") {  // CreateDiskShape, disk.cpp:163-173; Disk ctor disk.h:49-58
            IileSphere sp;
            std::memset(&sp, 0, sizeof(sp));
            std::memcpy(sp.objectToWorld, ctm.m.m, sizeof(float) * 16);
            std::memcpy(sp.worldToObject, ctm.inv.m, sizeof(float) * 16);
            auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
            sp.radius = ps.oneFloat("radius", 1.f);
            sp.zMin = ps.oneFloat("height", 0.f);        // height
            sp.zMax = ps.oneFloat("innerradius", 0.f);   // inner radius
            const float phimax = ps.oneFloat("phimax", 360.f);
            sp.phiMax = radians(clampf(phimax, 0, 360));
            sp.phiMaxDegrees = phimax;
            sp.flags = (gs.reverseOrientation ? 1u : 0u) | (ctm.swapsHandedness() ? 2u : 0u) | 4u;
            scene->spheres.push_back(sp);
            IilePrim pr;
            std::memset(&pr, 0, sizeof(pr));
            pr.type = IILE_PRIM_SPHERE;
            pr.v[0] = (int)scene->spheres.size() - 1;
            pr.material = material;
            pr.light = -1;
            addPrim(pr);
        } else if (name == "sphere") {  // CreateSphereShape, sphere.cpp:317-327; ctor sphere.h:50-59
            float radius = ps.oneFloat("radius", 1.f);
            float zmin = ps.oneFloat("zmin", -radius);
            float zmax = ps.oneFloat("zmax", radius);
            float phimax = ps.oneFloat("phimax", 360.f);
            auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
            IileSphere sp;
            std::memset(&sp, 0, sizeof(sp));
            std::memcpy(sp.objectToWorld, ctm.m.m, sizeof(float) * 16);
            std::memcpy(sp.worldToObject, ctm.inv.m, sizeof(float) * 16);
            sp.radius = radius;
            sp.zMin = clampf(std::min(zmin, zmax), -radius, radius);
            sp.zMax = clampf(std::max(zmin, zmax), -radius, radius);
            sp.thetaMin = std::acos(clampf(std::min(zmin, zmax) / radius, -1, 1));
            sp.thetaMax = std::acos(clampf(std::max(zmin, zmax) / radius, -1, 1));
            sp.phiMax = radians(clampf(phimax, 0, 360));
            sp.phiMaxDegrees = phimax;
            sp.flags = (gs.reverseOrientation ? 1u : 0u) | (ctm.swapsHandedness() ? 2u : 0u);
            scene->spheres.push_back(sp);
            IilePrim pr;
            std::memset(&pr, 0, sizeof(pr));
            pr.type = IILE_PRIM_SPHERE;
            pr.v[0] = (int)scene->spheres.size() - 1;
            pr.material = material;
            pr.light = -1;
            addPrim(pr);
        } else {
            fail("Shape \"" + name + "\" is not supported by the B200 path (trianglemesh/loopsubdiv/plymesh/sphere/disk/cylinder only)");
        }
    }

    // InfiniteAreaLight's ctor for a constant L (infinite.cpp:57-84): Lmap is a 1x1 MIPMap; the
    // sampling image is 2x2, img[u + 2v] = Lmap->Lookup((up, vp), fwidth).y() * sin(Pi (v+.5)/2), and
    // Distribution2D (sampling.cpp:159-171) over it.  MIPMap::Lookup at these widths is triangle(0, st)
    // (mipmap.h:244-274): a bilinear blend of the one texel with itself, kept term by term.
    static void envTriangle(const float *L, float s0, float t0, float *out) {
        float s = s0 * 1 - 0.5f, t = t0 * 1 - 0.5f;
        float fs = std::floor(s), ft = std::floor(t);
        float ds = s - (int)fs, dt = t - (int)ft;
        for (int c = 0; c < 3; ++c)
            out[c] = (1 - ds) * (1 - dt) * L[c] + (1 - ds) * dt * L[c] + ds * (1 - dt) * L[c] + ds * dt * L[c];
    }
    static void distribution1D(const float *f, int n, float *cdf, float *funcInt) {  // sampling.h:57-72
        cdf[0] = 0;
        for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / n;
        *funcInt = cdf[n];
        if (*funcInt == 0)
            for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
        else
            for (int i = 1; i < n + 1; ++i) cdf[i] /= *funcInt;
    }
    static void buildConstantEnvDistribution(IileLight *L) {
        const int width = 2, height = 2;
        const float fwidth = 0.5f / std::min(width, height);
        (void)fwidth;  // level = Log2(0.25) < 0 -> triangle(0, st)
        for (int v = 0; v < height; ++v) {
            float vp = (v + .5f) / (float)height;
            float sinTheta = std::sin(kPi * (v + .5f) / height);
            for (int u = 0; u < width; ++u) {
                float up = (u + .5f) / (float)width;
                float rgb[3];
                envTriangle(L->L, up, vp, rgb);
                float y = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
                L->envFunc[v][u] = y;
                L->envFunc[v][u] *= sinTheta;
            }
        }
        for (int v = 0; v < height; ++v) distribution1D(L->envFunc[v], width, L->envCdf[v], &L->envFuncInt[v]);
        for (int v = 0; v < height; ++v) L->envMargFunc[v] = L->envFuncInt[v];
        distribution1D(L->envMargFunc, height, L->envMargCdf, &L->envMargFuncInt);
    }

    void lightSource(const std::string &name, const ParamSet &ps) {
        if (!inWorld) fail("Scene description must be inside world block; \"LightSource\" not allowed.");
        if (name == "spot") {  // CreateSpotLight, spot.cpp:99-124
            float I[3] = {1, 1, 1}, sc[3] = {1, 1, 1};
            ps.oneSpectrum("I", I);
            ps.oneSpectrum("scale", sc);
            const float coneangle = ps.oneFloat("coneangle", 30.f), conedelta = ps.oneFloat("conedeltaangle", 5.f);
            V3 from(0, 0, 0), to(0, 0, 1);
            ps.onePoint3("from", &from);
            ps.onePoint3("to", &to);
            const V3 dir = normalize(to - from);
            // CoordinateSystem(dir, &du, &dv), geometry.h:1020-1027
            V3 du = std::abs(dir.x) > std::abs(dir.y)
                        ? divide(V3(-dir.z, 0, dir.x), std::sqrt(dir.x * dir.x + dir.z * dir.z))
                        : divide(V3(0, dir.z, -dir.y), std::sqrt(dir.y * dir.y + dir.z * dir.z));
            V3 dv = cross(dir, du);
            Xform dirToZ(M44(du.x, du.y, du.z, 0, dv.x, dv.y, dv.z, 0, dir.x, dir.y, dir.z, 0, 0, 0, 0, 1));
            Xform l2w = ctm * xfTranslate(from) * inverse(dirToZ);
            V3 p = l2w.point(V3(0, 0, 0));
            IileLight L;
            std::memset(&L, 0, sizeof(L));
            L.type = IILE_LIGHT_SPOT;
            L.prim = -1;
            L.pos[0] = p.x; L.pos[1] = p.y; L.pos[2] = p.z;
            for (int i = 0; i < 3; ++i) L.L[i] = I[i] * sc[i];
            std::memcpy(L.lightToWorld, l2w.m.m, sizeof(float) * 16);
            std::memcpy(L.worldToLight, l2w.inv.m, sizeof(float) * 16);
            L.coneAngle = coneangle;              // SpotLight(light2world, medium, I * sc, coneangle, coneangle - conedelta)
            L.falloffStart = coneangle - conedelta;
            L.cosTotalWidth = std::cos(radians(L.coneAngle));
            L.cosFalloffStart = std::cos(radians(L.falloffStart));
            scene->lights.push_back(L);
            return;
        }
        if (name == "infinite" || name == "exinfinite") {  // CreateInfiniteLight, infinite.cpp:176-186 (api.cpp:714-715)
            if (!ps.oneString("mapname", "").empty())
                fail("LightSource \"infinite\" with a \"mapname\" (environment map) is not supported by the B200 path");
            float Lr[3] = {1, 1, 1}, sc[3] = {1, 1, 1};
            ps.oneSpectrum("L", Lr);
            ps.oneSpectrum("scale", sc);
            IileLight L;
            std::memset(&L, 0, sizeof(L));
            L.type = IILE_LIGHT_INFINITE;
            L.prim = -1;
            for (int i = 0; i < 3; ++i) L.L[i] = Lr[i] * sc[i];
            std::memcpy(L.lightToWorld, ctm.m.m, sizeof(float) * 16);
            std::memcpy(L.worldToLight, ctm.inv.m, sizeof(float) * 16);
            buildConstantEnvDistribution(&L);
            scene->lights.push_back(L);  // worldRadius: once the scene bound is known (bvh_build.cpp)
            return;
        }
        if (name == "distant") {  // CreateDistantLight, distant.cpp:93-102
            float Lr[3] = {1, 1, 1}, sc[3] = {1, 1, 1};
            ps.oneSpectrum("L", Lr);
            ps.oneSpectrum("scale", sc);
            V3 from(0, 0, 0), to(0, 0, 1);
            ps.onePoint3("from", &from);
            ps.onePoint3("to", &to);
            const V3 dir = from - to;
            const V3 w = normalize(ctm.vector(dir));
            IileLight L;
            std::memset(&L, 0, sizeof(L));
            L.type = IILE_LIGHT_DISTANT;
            L.prim = -1;
            L.pos[0] = w.x; L.pos[1] = w.y; L.pos[2] = w.z;
            for (int i = 0; i < 3; ++i) L.L[i] = Lr[i] * sc[i];
            std::memcpy(L.lightToWorld, ctm.m.m, sizeof(float) * 16);
            std::memcpy(L.worldToLight, ctm.inv.m, sizeof(float) * 16);
            L.dir[0] = dir.x; L.dir[1] = dir.y; L.dir[2] = dir.z;
            scene->lights.push_back(L);  // worldRadius: once the scene bound is known (bvh_build.cpp)
            return;
        }
        if (name != "point")
            fail("LightSource \"" + name + "\" is not supported by the B200 path (point, spot, distant, infinite without a map, and "
                 "AreaLightSource diffuse, only)");
        // CreatePointLight, point.cpp:80-88; PointLight ctor point.h:52-56
        float I[3] = {1, 1, 1}, sc[3] = {1, 1, 1};
        ps.oneSpectrum("I", I);
        ps.oneSpectrum("scale", sc);
        V3 from(0, 0, 0);
        ps.onePoint3("from", &from);
        Xform l2w = xfTranslate(from) * ctm;
        V3 p = l2w.point(V3(0, 0, 0));
        IileLight L;
        std::memset(&L, 0, sizeof(L));
        L.type = IILE_LIGHT_POINT;
        L.prim = -1;
        L.pos[0] = p.x; L.pos[1] = p.y; L.pos[2] = p.z;
        for (int i = 0; i < 3; ++i) L.L[i] = I[i] * sc[i];
        scene->lights.push_back(L);
    }

    // CreateImageFloatTexture / CreateImageSpectrumTexture, imagemap.cpp:107-185
    TexRef imageMap(const ParamSet &ps, int channels) {
        const std::string mapping = ps.oneString("mapping", "uv");
        if (mapping != "uv") fail("2D texture mapping \"" + mapping + "\" is not supported by the B200 path (uv only)");
        IileTexture t;
        std::memset(&t, 0, sizeof(t));
        t.type = IILE_TEX_IMAGEMAP;
        t.channels = channels;
        t.tex1 = t.tex2 = -1;
        t.su = ps.oneFloat("uscale", 1.f);
        t.sv = ps.oneFloat("vscale", 1.f);
        t.du = ps.oneFloat("udelta", 0.f);
        t.dv = ps.oneFloat("vdelta", 0.f);
        t.maxAnisotropy = ps.oneFloat("maxanisotropy", 8.f);
        // the EWA filter walks every texel of the ellipse's bounding box (mipmap.h:396-414): its side grows with
        // the anisotropy, and a kernel thread must not be handed an unbounded loop
        if (!(t.maxAnisotropy >= 0.f && t.maxAnisotropy <= 64.f))
            fail("Texture \"imagemap\": \"maxanisotropy\" above 64 is not supported by the B200 path");
        t.trilinear = ps.oneBool("trilinear", false);
        const std::string wrap = ps.oneString("wrap", "repeat");
        t.wrap = wrap == "black" ? IILE_WRAP_BLACK : (wrap == "clamp" ? IILE_WRAP_CLAMP : IILE_WRAP_REPEAT);
        const float scale = ps.oneFloat("scale", 1.f);
        const std::string fn = ps.oneString("filename", "");
        if (fn.empty()) fail("Texture \"imagemap\": no \"filename\"");
        auto ext = [&](const char *e) {
            const size_t n = std::strlen(e);
            return fn.size() >= n && fn.compare(fn.size() - n, n, e) == 0;
        };
        const bool gamma = ps.oneBool("gamma", ext(".tga") || ext(".png"));
        t.image = loadImageMap(scene, resolve(fn), channels, t.wrap, scale, gamma);
        scene->textures.push_back(t);
        TexRef r;
        r.index = (int)scene->textures.size() - 1;
        return r;
    }
    // operators above the leaves of a texture reference (IILE_TEX_MAX_DEPTH bounds the device's evaluation)
    int texDepth(const TexRef &r) const {
        if (r.index < 0) return 0;
        const IileTexture &t = scene->textures[r.index];
        if (t.type != IILE_TEX_SCALE && t.type != IILE_TEX_MIX && t.type != IILE_TEX_CHECKERBOARD) return 0;
        TexRef a, b, c;
        a.index = t.tex1, b.index = t.tex2, c.index = t.type == IILE_TEX_MIX ? t.image : -1;
        return 1 + std::max(texDepth(a), std::max(texDepth(b), texDepth(c)));
    }
    TexRef pushOperator(IileTexture t, const char *what) {
        scene->textures.push_back(t);
        TexRef r;
        r.index = (int)scene->textures.size() - 1;
        if (texDepth(r) > IILE_TEX_MAX_DEPTH)
            fail(std::string("Texture \"") + what + "\": scale / mix textures nested more than " + std::to_string(IILE_TEX_MAX_DEPTH) +
                 " deep are not supported by the B200 path");
        return r;
    }
    // ScaleTexture (scale.h:48-64) of two references
    TexRef scaleOf(const TexRef &a, const TexRef &b, int channels) {
        TexRef r;
        if (a.index < 0 && b.index < 0) {
            for (int i = 0; i < 3; ++i) r.v[i] = a.v[i] * b.v[i];
            return r;
        }
        IileTexture t;
        std::memset(&t, 0, sizeof(t));
        t.type = IILE_TEX_SCALE;
        t.channels = channels;
        t.image = -1;
        t.tex1 = textureIndex(a, channels);
        t.tex2 = textureIndex(b, channels);
        return pushOperator(t, "scale");
    }
    // MixTexture (mix.h:49-66): (1 - amount) * tex1 + amount * tex2, the amount a float texture
    TexRef mixOf(const TexRef &a, const TexRef &b, const TexRef &amount, int channels) {
        TexRef r;
        if (a.index < 0 && b.index < 0 && amount.index < 0) {
            for (int i = 0; i < 3; ++i) r.v[i] = (1 - amount.v[0]) * a.v[i] + amount.v[0] * b.v[i];
            return r;
        }
        IileTexture t;
        std::memset(&t, 0, sizeof(t));
        t.type = IILE_TEX_MIX;
        t.channels = channels;
        t.tex1 = textureIndex(a, channels);
        t.tex2 = textureIndex(b, channels);
        t.image = amount.index;  // -1: the constant below
        t.value[0] = amount.v[0];
        return pushOperator(t, "mix");
    }
    void texture(const std::string &name, const std::string &type, const std::string &cls,
                 const ParamSet &ps) {
        // pbrtTexture, api.cpp:1231-1280: constant, scale, imagemap; mix of constants
        const bool isFloat = type == "float";
        if (!isFloat && type != "spectrum" && type != "color") fail("Texture type \"" + type + "\" unknown.");
        const int channels = isFloat ? 1 : 3;
        auto ref = [&](const char *n, float def) {
            return isFloat ? floatRef(ps, ParamSet(), n, def) : spectrumRef(ps, ParamSet(), n, def);
        };
        TexRef v;
        if (cls == "constant") {
            v = ref("value", 1.f);
            if (v.index >= 0) fail("Texture \"constant\": \"value\" has to be a constant");
        } else if (cls == "scale")
            v = scaleOf(ref("tex1", 1.f), ref("tex2", 1.f), channels);
        else if (cls == "mix")  // CreateMixFloatTexture / CreateMixSpectrumTexture, mix.cpp:39-53
            v = mixOf(ref("tex1", 0.f), ref("tex2", 1.f), floatRef(ps, ParamSet(), "amount", 0.5f), channels);
        else if (cls == "checkerboard") {  // CreateCheckerboard{Float,Spectrum}Texture, checkerboard.cpp:39-160
            if (ps.oneInt("dimension", 2) != 2) fail("Texture \"checkerboard\": only \"dimension\" 2 is supported by the B200 path");
            if (ps.oneString("mapping", "uv") != "uv") fail("2D texture mapping \"" + ps.oneString("mapping", "uv") + "\" is not supported by the B200 path (uv only)");
            IileTexture t;
            std::memset(&t, 0, sizeof(t));
            t.type = IILE_TEX_CHECKERBOARD;
            t.channels = channels;
            t.image = -1;
            t.tex1 = textureIndex(ref("tex1", 1.f), channels);
            t.tex2 = textureIndex(ref("tex2", 0.f), channels);
            t.su = ps.oneFloat("uscale", 1.f);
            t.sv = ps.oneFloat("vscale", 1.f);
            t.du = ps.oneFloat("udelta", 0.f);
            t.dv = ps.oneFloat("vdelta", 0.f);
            t.trilinear = ps.oneString("aamode", "closedform") == "none";  // anything else: closedform, with the reference's warning
            v = pushOperator(t, "checkerboard");
        } else if (cls == "imagemap")
            v = imageMap(ps, channels);
        else
            fail("Texture class \"" + cls + "\" is not supported by the B200 path (constant / scale / mix / checkerboard / imagemap)");
        (isFloat ? gs.floatTextures : gs.spectrumTextures)[name] = v;
    }

    // ---- option block -> IileRenderSetup (MakeCamera/MakeFilm/MakeFilter/MakeSampler/
    //      CreatePathIntegrator/CreateBVHAccelerator) -------------------------------
    static int roundUpPow2(int v) {  // pbrt.h:362-370
        int p = 1;
        while (p < v) p <<= 1;
        return p;
    }
    void finishSetup() {
        IileRenderSetup &S = scene->setup;
        defaultRenderSetup(&S);
        if (cameraName != "perspective" && cameraName != "orthographic" && cameraName != "environment")
            fail("Camera \"" + cameraName + "\" is not supported by the B200 path (perspective, orthographic and environment only)");
        if (filmName != "image") fail("Film \"" + filmName + "\" unknown.");
        if (integratorName != "path" && integratorName != "directlighting" && integratorName != "whitted")
            fail("Integrator \"" + integratorName + "\" is not supported by the B200 path (path, directlighting and whitted only)");
        if (accelName != "bvh") fail("Accelerator \"" + accelName + "\" is not supported by the B200 path (bvh only)");
        if (samplerName != "random" && samplerName != "halton" && samplerName != "sobol")
            fail("Sampler \"" + samplerName + "\" is not supported by the B200 path (random, halton and sobol only)");
        // film.cpp:259-303
        S.xRes = filmParams.oneInt("xresolution", 1280);
        S.yRes = filmParams.oneInt("yresolution", 720);
        std::string fn = filmParams.oneString("filename", "pbrt.exr");
        std::snprintf(S.filename, sizeof(S.filename), "%s", fn.c_str());
        auto clamp01 = [](float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); };
        if (const Param *cr = filmParams.floats("cropwindow")) {
            if (cr->nums.size() != 4) fail("\"cropwindow\" expects 4 values");
            float c0 = (float)cr->nums[0], c1 = (float)cr->nums[1], c2 = (float)cr->nums[2], c3 = (float)cr->nums[3];
            S.cropWindow[0] = clamp01(std::min(c0, c1));
            S.cropWindow[1] = clamp01(std::max(c0, c1));
            S.cropWindow[2] = clamp01(std::min(c2, c3));
            S.cropWindow[3] = clamp01(std::max(c2, c3));
        }
        S.filmScale = filmParams.oneFloat("scale", 1.f);
        S.diagonal = filmParams.oneFloat("diagonal", 35.f);
        S.maxSampleLuminance = filmParams.oneFloat("maxsampleluminance", kInfinity);
        // filters/*.cpp Create*Filter
        if (filterName == "box") {
            S.filterType = IILE_FILTER_BOX;
            S.filterRadius[0] = filterParams.oneFloat("xwidth", 0.5f);
            S.filterRadius[1] = filterParams.oneFloat("ywidth", 0.5f);
        } else if (filterName == "gaussian") {
            S.filterType = IILE_FILTER_GAUSSIAN;
            S.filterRadius[0] = filterParams.oneFloat("xwidth", 2.f);
            S.filterRadius[1] = filterParams.oneFloat("ywidth", 2.f);
            S.filterParam[0] = filterParams.oneFloat("alpha", 2.f);
        } else if (filterName == "mitchell") {
            S.filterType = IILE_FILTER_MITCHELL;
            S.filterRadius[0] = filterParams.oneFloat("xwidth", 2.f);
            S.filterRadius[1] = filterParams.oneFloat("ywidth", 2.f);
            S.filterParam[0] = filterParams.oneFloat("B", 1.f / 3.f);
            S.filterParam[1] = filterParams.oneFloat("C", 1.f / 3.f);
        } else if (filterName == "sinc") {
            S.filterType = IILE_FILTER_SINC;
            S.filterRadius[0] = filterParams.oneFloat("xwidth", 4.f);
            S.filterRadius[1] = filterParams.oneFloat("ywidth", 4.f);
            S.filterParam[0] = filterParams.oneFloat("tau", 3.f);
        } else if (filterName == "triangle") {
            S.filterType = IILE_FILTER_TRIANGLE;
            S.filterRadius[0] = filterParams.oneFloat("xwidth", 2.f);
            S.filterRadius[1] = filterParams.oneFloat("ywidth", 2.f);
        } else
            fail("Filter \"" + filterName + "\" unknown.");
        // perspective.cpp:293-329
        std::memcpy(S.cameraToWorld, cameraToWorld.m.m, sizeof(float) * 16);
        std::memcpy(S.worldToCamera, cameraToWorld.inv.m, sizeof(float) * 16);
        S.shutterOpen = cameraParams.oneFloat("shutteropen", 0.f);
        S.shutterClose = cameraParams.oneFloat("shutterclose", 1.f);
        if (S.shutterClose < S.shutterOpen) std::swap(S.shutterClose, S.shutterOpen);
        S.lensRadius = cameraParams.oneFloat("lensradius", 0.f);
        S.focalDistance = cameraParams.oneFloat("focaldistance", 1e6f);
        float frame = cameraParams.oneFloat("frameaspectratio", float(S.xRes) / float(S.yRes));
        if (frame > 1.f) {
            S.screenWindow[0] = -frame; S.screenWindow[1] = frame;
            S.screenWindow[2] = -1.f;   S.screenWindow[3] = 1.f;
        } else {
            S.screenWindow[0] = -1.f;          S.screenWindow[1] = 1.f;
            S.screenWindow[2] = -1.f / frame;  S.screenWindow[3] = 1.f / frame;
        }
        if (const Param *sw = cameraParams.floats("screenwindow")) {
            if (sw->nums.size() != 4) fail("\"screenwindow\" should have four values");
            for (int i = 0; i < 4; ++i) S.screenWindow[i] = (float)sw->nums[i];
        }
        S.fov = cameraParams.oneFloat("fov", 90.f);
        float halffov = cameraParams.oneFloat("halffov", -1.f);
        if (halffov > 0.f) S.fov = 2.f * halffov;
        if (cameraName == "orthographic") S.cameraType = IILE_CAMERA_ORTHOGRAPHIC;  // orthographic.cpp:120-168: same parameters, no fov
        if (cameraName == "environment") {  // environment.cpp:59-100: no lens; the screen window is parsed but unused
            S.cameraType = IILE_CAMERA_ENVIRONMENT;
            S.lensRadius = 0.f;
        }
        if (samplerName == "halton") {  // halton.cpp:124-130
            S.samplerType = IILE_SAMPLER_HALTON;
            S.pixelSamples = samplerParams.oneInt("pixelsamples", 16);
            S.samplePixelCenter = samplerParams.oneBool("samplepixelcenter", false);
        } else if (samplerName == "sobol") {  // CreateSobolSampler, sobol.cpp:64-70; the ctor rounds up to a power of two (sobol.h:64-70)
            S.samplerType = IILE_SAMPLER_SOBOL;
            S.pixelSamples = roundUpPow2(samplerParams.oneInt("pixelsamples", 16));
        } else {  // random.cpp:74-77
            S.samplerType = IILE_SAMPLER_RANDOM;
            S.pixelSamples = samplerParams.oneInt("pixelsamples", 4);
        }
        // IILE: CreatePathIntegrator swaps in a SobolSampler when IILE_PATH_SAMPLES_OVERRIDE is set (path.cpp:201-212)
        if (integratorName == "path")
            if (const char *ov = std::getenv("IILE_PATH_SAMPLES_OVERRIDE")) {
                const int spp = std::atoi(ov);
                if (spp <= 0) fail("IILE_PATH_SAMPLES_OVERRIDE must be a positive sample count");
                S.samplerType = IILE_SAMPLER_SOBOL;
                S.pixelSamples = roundUpPow2(spp);
            }
        if (integratorName == "directlighting") {  // CreateDirectLightingIntegrator, directlighting.cpp:101-132
            S.integratorType = IILE_INTEGRATOR_DIRECT;
            std::string st = integratorParams.oneString("strategy", "all");
            // an unknown strategy is a Warning and "all" in the reference (directlighting.cpp:111-116)
            S.directStrategy = st == "one" ? IILE_DIRECT_ONE : IILE_DIRECT_ALL;
        }
        if (integratorName == "whitted") {  // CreateWhittedIntegrator, whitted.cpp:90-109
            S.integratorType = IILE_INTEGRATOR_WHITTED;
        }
        // path.cpp:214-231 / directlighting.cpp:104-110 / whitted.cpp:93-106
        S.maxDepth = integratorParams.oneInt("maxdepth", 5);
        S.rrThreshold = integratorParams.oneFloat("rrthreshold", 1.f);
        if (const Param *pb = integratorParams.ints("pixelbounds")) {
            if (pb->nums.size() != 4) fail("Expected four values for \"pixelbounds\" parameter.");
            S.hasPixelBounds = 1;
            for (int i = 0; i < 4; ++i) S.pixelBounds[i] = (int)pb->nums[i];
        }
        std::string ls = integratorParams.oneString("lightsamplestrategy", "spatial");
        if (ls == "uniform") S.lightSampleStrategy = IILE_LIGHTSTRATEGY_UNIFORM;
        else if (ls == "power") S.lightSampleStrategy = IILE_LIGHTSTRATEGY_POWER;
        else if (ls == "spatial") S.lightSampleStrategy = IILE_LIGHTSTRATEGY_SPATIAL;
        else fail("Light sample distribution type \"" + ls + "\" unknown.");
        // bvh.cpp:740-760
        std::string sm = accelParams.oneString("splitmethod", "sah");
        if (sm == "sah") S.splitMethod = IILE_SPLIT_SAH;
        else if (sm == "middle") S.splitMethod = IILE_SPLIT_MIDDLE;
        else if (sm == "equal") S.splitMethod = IILE_SPLIT_EQUAL;
        else if (sm == "hlbvh") S.splitMethod = IILE_SPLIT_HLBVH;
        else S.splitMethod = IILE_SPLIT_SAH;  // reference warns and falls back to sah
        S.maxNodePrims = accelParams.oneInt("maxnodeprims", 4);
    }

    // ---- directive loop ------------------------------------------------------
    void parseFile(const std::string &path, int depth) {
        if (depth > 32) fail("Include nesting too deep");
        Tokenizer tz(path);
        std::string tok;
        bool haveUnget = false;
        std::string ungot;
        auto next = [&](bool required) -> bool {
            if (haveUnget) { tok = ungot; haveUnget = false; return true; }
            if (!tz.next(&tok)) {
                if (required) tz.err("premature EOF");
                return false;
            }
            return true;
        };
        auto unget = [&]() { ungot = tok; haveUnget = true; };
        auto quoted = [&]() -> std::string {
            next(true);
            if (!isQuoted(tok)) tz.err("\"" + tok + "\": expected quoted string");
            return tok.substr(1, tok.size() - 2);
        };
        auto number = [&]() -> float {
            next(true);
            return (float)parseNumber(tz, tok);
        };
        // parser.cpp:455-560 parseParamList
        auto params = [&]() -> ParamSet {
            ParamSet ps;
            while (true) {
                if (!next(false)) return ps;
                if (!isQuoted(tok)) { unget(); return ps; }
                Param p;
                std::string why;
                std::string decl = tok.substr(1, tok.size() - 2);
                if (!lookupType(decl, &p.type, &p.name, &why)) tz.err(why);
                bool stringy = (p.type == PT::String || p.type == PT::Texture || p.type == PT::Bool);
                auto addValue = [&](const std::string &v) {
                    if (isQuoted(v)) {
                        if (!stringy) tz.err("Expected numeric parameter value for \"" + decl + "\"");
                        p.strs.push_back(v.substr(1, v.size() - 2));
                    } else if (p.type == PT::Bool && (v == "true" || v == "false")) {
                        p.strs.push_back(v);
                    } else {
                        if (stringy) tz.err("Expected string parameter value for \"" + decl + "\"");
                        p.nums.push_back(parseNumber(tz, v));
                    }
                };
                next(true);
                if (tok == "[") {
                    while (true) {
                        next(true);
                        if (tok == "]") break;
                        addValue(tok);
                    }
                } else
                    addValue(tok);
                if (p.type == PT::Bool)
                    for (const std::string &b : p.strs)
                        if (b != "true" && b != "false")
                            tz.err("Value \"" + b + "\" unknown for Boolean parameter \"" + p.name + "\"");
                if (p.type == PT::Int)
                    for (double &v : p.nums) v = (double)(int)v;
                ps.items.push_back(std::move(p));
            }
        };
        auto applyXf = [&](const Xform &t) { ctm = ctm * t; };

        while (next(false)) {
            const std::string d = tok;
            if (sawWorldEnd) tz.err("directive \"" + d + "\" after WorldEnd is not supported by the B200 path");
            if (d == "AttributeBegin") {
                gsStack.push_back(gs); xfStack.push_back(ctm); pushKind.push_back('a');
            } else if (d == "AttributeEnd") {
                if (gsStack.empty() || pushKind.back() != 'a') tz.err("Unmatched AttributeEnd encountered.");
                gs = gsStack.back(); gsStack.pop_back();
                ctm = xfStack.back(); xfStack.pop_back(); pushKind.pop_back();
            } else if (d == "TransformBegin") {
                xfStack.push_back(ctm); pushKind.push_back('t');
            } else if (d == "TransformEnd") {
                if (xfStack.empty() || pushKind.back() != 't') tz.err("Unmatched TransformEnd encountered.");
                ctm = xfStack.back(); xfStack.pop_back(); pushKind.pop_back();
            } else if (d == "Identity") {
                ctm = Xform();
            } else if (d == "Translate") {
                float x = number(), y = number(), z = number();
                applyXf(xfTranslate(V3(x, y, z)));
            } else if (d == "Scale") {
                float x = number(), y = number(), z = number();
                applyXf(xfScale(x, y, z));
            } else if (d == "Rotate") {
                float a = number(), x = number(), y = number(), z = number();
                applyXf(xfRotate(a, V3(x, y, z)));
            } else if (d == "LookAt") {
                float v[9];
                for (float &f : v) f = number();
                applyXf(xfLookAt(V3(v[0], v[1], v[2]), V3(v[3], v[4], v[5]), V3(v[6], v[7], v[8])));
            } else if (d == "ConcatTransform" || d == "Transform") {
                next(true);
                if (tok != "[") tz.err("expected \"[\"");
                float v[16];
                for (float &f : v) f = number();
                next(true);
                if (tok != "]") tz.err("expected \"]\"");
                // api.cpp:972-1000: the file stores the matrix column-major
                M44 m(v[0], v[4], v[8], v[12], v[1], v[5], v[9], v[13], v[2], v[6], v[10], v[14],
                      v[3], v[7], v[11], v[15]);
                if (d == "Transform") ctm = Xform(m);
                else applyXf(Xform(m));
            } else if (d == "CoordinateSystem") {
                namedCS[quoted()] = ctm;
            } else if (d == "CoordSysTransform") {
                std::string n = quoted();
                auto it = namedCS.find(n);
                if (it != namedCS.end()) ctm = it->second;  // reference warns when missing
            } else if (d == "ReverseOrientation") {
                gs.reverseOrientation = !gs.reverseOrientation;
            } else if (d == "Camera") {
                cameraName = quoted(); cameraParams = params();
                cameraToWorld = inverse(ctm);
                namedCS["camera"] = cameraToWorld;
            } else if (d == "Film") {
                filmName = quoted(); filmParams = params();
            } else if (d == "Sampler") {
                samplerName = quoted(); samplerParams = params();
            } else if (d == "PixelFilter") {
                filterName = quoted(); filterParams = params();
            } else if (d == "Integrator") {
                integratorName = quoted(); integratorParams = params();
            } else if (d == "Accelerator") {
                accelName = quoted(); accelParams = params();
            } else if (d == "WorldBegin") {
                inWorld = true;
                ctm = Xform();
                namedCS["world"] = ctm;
            } else if (d == "WorldEnd") {
                if (!gsStack.empty() || !xfStack.empty()) tz.err("Missing end to AttributeBegin / TransformBegin");
                sawWorldEnd = true;
            } else if (d == "Material") {
                auto m = std::make_shared<MaterialDef>();
                m->type = quoted();
                m->params = params();
                gs.material = m;
            } else if (d == "MakeNamedMaterial") {
                std::string n = quoted();
                auto m = std::make_shared<MaterialDef>();
                m->params = params();
                m->type = m->params.oneString("type", "");
                if (m->type.empty()) tz.err("No parameter string \"type\" found in MakeNamedMaterial");
                gs.namedMaterials[n] = m;
            } else if (d == "NamedMaterial") {
                std::string n = quoted();
                auto it = gs.namedMaterials.find(n);
                if (it == gs.namedMaterials.end()) tz.err("NamedMaterial \"" + n + "\" unknown.");
                gs.material = it->second;
            } else if (d == "Texture") {
                std::string n = quoted(), ty = quoted(), cls = quoted();
                ParamSet ps = params();
                texture(n, ty, cls, ps);
            } else if (d == "LightSource") {
                std::string n = quoted();
                ParamSet ps = params();
                lightSource(n, ps);
            } else if (d == "AreaLightSource") {
                gs.areaLight = quoted();
                gs.areaLightParams = params();
            } else if (d == "Shape") {
                std::string n = quoted();
                ParamSet ps = params();
                shape(n, ps);
            } else if (d == "Include") {
                std::string f = quoted();
                parseFile(resolve(f), depth + 1);
            } else if (d == "TransformTimes") {
                number(); number();
            } else if (d == "ActiveTransform") {
                next(true);
                if (tok != "All") tz.err("ActiveTransform " + tok + " (motion blur) is not supported by the B200 path");
            } else if (d == "ObjectBegin" || d == "ObjectEnd" || d == "ObjectInstance") {
                tz.err("object instancing (" + d + ") is not supported by the B200 path");
            } else if (d == "MakeNamedMedium" || d == "MediumInterface") {
                tz.err("participating media (" + d + ") are not supported by the B200 path");
            } else {
                tz.err("unknown directive \"" + d + "\"");
            }
        }
    }
};

}  // namespace

void defaultRenderSetup(IileRenderSetup *s) {
    std::memset(s, 0, sizeof(*s));
    M44 id;
    std::memcpy(s->cameraToWorld, id.m, sizeof(float) * 16);
    std::memcpy(s->worldToCamera, id.m, sizeof(float) * 16);
    s->fov = 90.f;
    s->focalDistance = 1e6f;
    s->screenWindow[0] = -1; s->screenWindow[1] = 1; s->screenWindow[2] = -1; s->screenWindow[3] = 1;
    s->shutterClose = 1.f;
    s->xRes = 1280; s->yRes = 720;
    s->cropWindow[0] = 0; s->cropWindow[1] = 1; s->cropWindow[2] = 0; s->cropWindow[3] = 1;
    s->filmScale = 1.f;
    s->maxSampleLuminance = kInfinity;
    s->diagonal = 35.f;
    s->filterType = IILE_FILTER_BOX;
    s->filterRadius[0] = s->filterRadius[1] = 0.5f;
    s->pixelSamples = 4;
    s->maxDepth = 5;
    s->rrThreshold = 1.f;
    s->lightSampleStrategy = IILE_LIGHTSTRATEGY_SPATIAL;
    s->maxNodePrims = 4;
    s->splitMethod = IILE_SPLIT_SAH;
    std::snprintf(s->filename, sizeof(s->filename), "pbrt.exr");
}

void loadPbrtFile(const std::string &path, HostScene *scene) {
    Loader L;
    L.scene = scene;
    size_t slash = path.find_last_of('/');
    L.searchDir = slash == std::string::npos ? "" : path.substr(0, slash);
    L.parseFile(path, 0);
    if (!L.sawWorldEnd) fail(path + ": no WorldEnd");
    L.finishSetup();
    if (scene->prims.empty()) fail(path + ": scene has no shapes");
    validateSetup(scene->setup);  // backstop: the directives' own checks name the parameter at fault
    buildBVH(scene);
}

}  // namespace iile
