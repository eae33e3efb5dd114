// oracle/ref_harness — TEST INFRASTRUCTURE.  Our own driver code that rebuilds
// a flat scene (the product's IILEFLT1 container, include/iile_b200.h structs)
// out of the REFERENCE's own classes, through their public constructors, and
// runs the reference's BVHAccel / PathIntegrator on it.  Linked against
// oracle/_ref/libpbrtref.a (the unmodified reference sources).  Nothing here
// is shipped or measured as the product; bench.py runs it only as the CPU
// baseline arm.
//
//   ref_harness bvh     <scene.flat> <nodes.bin> <order.bin>
//   ref_harness raycast <scene.flat> <rays.bin> <out.bin> closest|any [nthreads]
//   ref_harness render  <scene.flat> <out.pfm> stock|pixelstream [nthreads] [spp] [sampleOffset]
//   ref_harness genrays <scene.flat> <nrays> <rays.bin>      (cfg-2 ray set)
//
// The BVH's private arrays are read through `#define private public` on the
// one header that declares them; LinearBVHNode is re-declared with the layout
// of bvh.cpp:95-104.
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

// every header bvh.h pulls in is included first, so that only BVHAccel's own
// declaration is seen with its private section opened
#include <atomic>
#include <iostream>
#include <memory>
#include <sstream>

#include "pbrt.h"
#include "primitive.h"
#define private public
#include "accelerators/bvh.h"
#undef private

#include "api.h"
#include "camera.h"
#include "cameras/environment.h"
#include "cameras/orthographic.h"
#include "cameras/perspective.h"
#include "film.h"
#include "filters/box.h"
#include "filters/gaussian.h"
#include "filters/mitchell.h"
#include "filters/sinc.h"
#include "filters/triangle.h"
#include "integrators/directlighting.h"
#include "integrators/path.h"
#include "integrators/whitted.h"
#include "lights/diffuse.h"
#include "lights/distant.h"
#include "lights/infinite.h"
#include "lights/point.h"
#include "lights/spot.h"
#include "materials/glass.h"
#include "materials/matte.h"
#include "materials/mirror.h"
#include "materials/plastic.h"
#include "materials/substrate.h"
#include "materials/translucent.h"
#include "materials/uber.h"
#include "parallel.h"
#include "rng.h"
#include "sampler.h"
#include "samplers/random.h"
#include "samplers/sobol.h"
#include "samplers/halton.h"
#include "sampling.h"
#include "scene.h"
#include "shapes/cylinder.h"
#include "shapes/disk.h"
#include "shapes/sphere.h"
#include "shapes/triangle.h"
#include "stats.h"
#include "textures/constant.h"

#include "harness.h"

using namespace pbrt;

namespace pbrt {
struct LinearBVHNode {  // bvh.cpp:95-104
    Bounds3f bounds;
    union {
        int primitivesOffset;
        int secondChildOffset;
    };
    uint16_t nPrimitives;
    uint8_t axis;
    uint8_t pad[1];
};
}  // namespace pbrt
static_assert(sizeof(pbrt::LinearBVHNode) == 32, "LinearBVHNode layout");

// ---- flat scene (struct Flat: harness.h) ------------------------------------
void die(const std::string &m) {
    fprintf(stderr, "ref_harness: %s\n", m.c_str());
    fflush(stderr);
    _exit(2);  // worker threads may be parked; skip static destructors
}

Flat loadFlat(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) die("cannot open " + path);
    char magic[8];
    f.read(magic, 8);
    const bool v1 = !std::memcmp(magic, "IILEFLT1", 8);  // 64-byte materials (ABI 1)
    if (!v1 && std::memcmp(magic, "IILEFLT2", 8)) die(path + ": not a flat scene");
    uint32_t c[8];
    f.read((char *)c, sizeof(c));
    Flat s;
    f.read((char *)&s.setup, sizeof(s.setup));
    s.P.resize((size_t)c[0] * 3);
    f.read((char *)s.P.data(), s.P.size() * 4);
    s.N.resize(c[1] ? (size_t)c[0] * 3 : 0);
    f.read((char *)s.N.data(), s.N.size() * 4);
    s.UV.resize(c[2] ? (size_t)c[0] * 2 : 0);
    f.read((char *)s.UV.data(), s.UV.size() * 4);
    s.prims.resize(c[3]);
    f.read((char *)s.prims.data(), s.prims.size() * sizeof(IilePrim));
    s.spheres.resize(c[4]);
    f.read((char *)s.spheres.data(), s.spheres.size() * sizeof(IileSphere));
    s.materials.resize(c[5]);
    for (IileMaterial &m : s.materials) {
        std::memset(&m, 0, sizeof(m));
        f.read((char *)&m, v1 ? 64 : sizeof(IileMaterial));
    }
    s.lights.resize(c[6]);
    for (IileLight &l : s.lights) {
        std::memset(&l, 0, sizeof(l));
        f.read((char *)&l, v1 ? 40 : sizeof(IileLight));
    }
    if (!f) die(path + ": truncated");
    return s;
}

// ---- reference objects built from the flat scene ----------------------------
static Transform *keep(RefScene &rs, const Transform &t) {
    rs.xforms.emplace_back(new Transform(t));
    return rs.xforms.back().get();
}

static Matrix4x4 mat(const float *m) {
    return Matrix4x4(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11],
                     m[12], m[13], m[14], m[15]);
}

static std::shared_ptr<Material> makeMaterial(const IileMaterial &m) {
    auto spec = [](const float *c) {
        Float rgb[3] = {c[0], c[1], c[2]};
        return std::shared_ptr<Texture<Spectrum>>(new ConstantTexture<Spectrum>(Spectrum::FromRGB(rgb)));
    };
    auto flt = [](float v) { return std::shared_ptr<Texture<Float>>(new ConstantTexture<Float>(v)); };
    std::shared_ptr<Texture<Float>> none;
    switch (m.type) {
    case IILE_MAT_MATTE:
        return std::make_shared<MatteMaterial>(spec(m.Kd), flt(m.sigma), none);
    case IILE_MAT_PLASTIC:
        return std::make_shared<PlasticMaterial>(spec(m.Kd), spec(m.Ks), flt(m.roughness), none,
                                                 m.remapRoughness != 0);
    case IILE_MAT_MIRROR:
        return std::make_shared<MirrorMaterial>(spec(m.Ks), none);
    case IILE_MAT_GLASS:
        return std::make_shared<GlassMaterial>(spec(m.Ks), spec(m.Kt), flt(m.uRoughness),
                                               flt(m.vRoughness), flt(m.eta), none,
                                               m.remapRoughness != 0);
    case IILE_MAT_SUBSTRATE:
        return std::make_shared<SubstrateMaterial>(spec(m.Kd), spec(m.Ks), flt(m.uRoughness), flt(m.vRoughness), none,
                                                   m.remapRoughness != 0);
    case IILE_MAT_TRANSLUCENT:
        return std::make_shared<TranslucentMaterial>(spec(m.Kd), spec(m.Ks), flt(m.roughness), spec(m.Kr), spec(m.Kt),
                                                     none, m.remapRoughness != 0);
    case IILE_MAT_UBER:
        return std::make_shared<UberMaterial>(spec(m.Kd), spec(m.Ks), spec(m.Kr), spec(m.Kt), flt(m.roughness),
                                              flt(m.uRoughness), flt(m.vRoughness), spec(m.opacity), flt(m.eta),
                                              none, m.remapRoughness != 0);
    }
    die("unknown material type");
    return nullptr;
}

// deindex=true gives every triangle its own three vertices with uv (0,0),(1,0),(0,1), so
// that isect.uv == (b1,b2) exactly; positions (and hence the tree and all hit tests) are
// unchanged.
void buildRefScene(const Flat &fs, RefScene &rs, bool deindex) {
    Transform *identity = keep(rs, Transform());
    std::vector<std::shared_ptr<Material>> mats;
    for (const IileMaterial &m : fs.materials) mats.push_back(makeMaterial(m));
    rs.prims.resize(fs.prims.size());
    std::vector<std::shared_ptr<AreaLight>> areaOf(fs.prims.size());
    std::vector<std::shared_ptr<Shape>> shapes(fs.prims.size());

    // group consecutive triangles with identical flags into one TriangleMesh
    size_t i = 0;
    while (i < fs.prims.size()) {
        const IilePrim &p = fs.prims[i];
        if (p.type == IILE_PRIM_SPHERE) {
            const IileSphere &sp = fs.spheres[p.v[0]];
            Transform *o2w = keep(rs, Transform(mat(sp.objectToWorld), mat(sp.worldToObject)));
            Transform *w2o = keep(rs, Transform(mat(sp.worldToObject), mat(sp.objectToWorld)));
            if (sp.flags & 8u)
                shapes[i] = std::make_shared<Cylinder>(o2w, w2o, (sp.flags & 1) != 0, sp.radius, sp.zMin, sp.zMax,
                                                       sp.phiMaxDegrees);
            else if (sp.flags & 4u)  // Shape "disk": height in zMin, inner radius in zMax
                shapes[i] = std::make_shared<Disk>(o2w, w2o, (sp.flags & 1) != 0, sp.zMin, sp.radius, sp.zMax,
                                                   sp.phiMaxDegrees);
            else
                shapes[i] = std::make_shared<Sphere>(o2w, w2o, (sp.flags & 1) != 0, sp.radius, sp.zMin,
                                                     sp.zMax, sp.phiMaxDegrees);
            ++i;
            continue;
        }
        size_t j = i;
        while (j < fs.prims.size() && fs.prims[j].type == IILE_PRIM_TRIANGLE &&
               fs.prims[j].flags == p.flags)
            ++j;
        const bool hasN = p.flags & IILE_PRIM_HAS_N, hasUV = p.flags & IILE_PRIM_HAS_UV;
        std::vector<int> idx;
        std::vector<Point3f> P;
        std::vector<Normal3f> N;
        std::vector<Point2f> UV;
        std::map<int, int> remap;
        for (size_t k = i; k < j; ++k)
            for (int c = 0; c < 3; ++c) {
                int v = fs.prims[k].v[c];
                int local;
                auto it = remap.find(v);
                if (deindex || it == remap.end()) {
                    local = (int)P.size();
                    remap[v] = local;
                    P.push_back(Point3f(fs.P[3 * v], fs.P[3 * v + 1], fs.P[3 * v + 2]));
                    if (hasN) N.push_back(Normal3f(fs.N[3 * v], fs.N[3 * v + 1], fs.N[3 * v + 2]));
                    if (deindex) UV.push_back(c == 0 ? Point2f(0, 0) : (c == 1 ? Point2f(1, 0) : Point2f(0, 1)));
                    else if (hasUV) UV.push_back(Point2f(fs.UV[2 * v], fs.UV[2 * v + 1]));
                } else
                    local = it->second;
                idx.push_back(local);
            }
        std::vector<std::shared_ptr<Shape>> tris = CreateTriangleMesh(
            identity, identity, (p.flags & IILE_PRIM_FLIP_NORMAL) != 0, (int)(j - i), idx.data(),
            (int)P.size(), P.data(), nullptr, hasN ? N.data() : nullptr,
            (deindex || hasUV) ? UV.data() : nullptr, nullptr, nullptr, nullptr);
        for (size_t k = i; k < j; ++k) shapes[k] = tris[k - i];
        i = j;
    }
    // lights, in scene order
    for (const IileLight &l : fs.lights) {
        Float rgb[3] = {l.L[0], l.L[1], l.L[2]};
        if (l.type == IILE_LIGHT_POINT) {
            rs.lights.push_back(std::make_shared<PointLight>(
                Translate(Vector3f(l.pos[0], l.pos[1], l.pos[2])), MediumInterface(), Spectrum::FromRGB(rgb)));
        } else if (l.type == IILE_LIGHT_SPOT) {
            rs.lights.push_back(std::make_shared<SpotLight>(Transform(mat(l.lightToWorld), mat(l.worldToLight)),
                                                            MediumInterface(), Spectrum::FromRGB(rgb), l.coneAngle,
                                                            l.falloffStart));
        } else if (l.type == IILE_LIGHT_INFINITE) {
            rs.lights.push_back(std::make_shared<InfiniteAreaLight>(Transform(mat(l.lightToWorld), mat(l.worldToLight)),
                                                                    Spectrum::FromRGB(rgb), 1, ""));
        } else if (l.type == IILE_LIGHT_DISTANT) {
            rs.lights.push_back(std::make_shared<DistantLight>(Transform(mat(l.lightToWorld), mat(l.worldToLight)),
                                                               Spectrum::FromRGB(rgb),
                                                               Vector3f(l.dir[0], l.dir[1], l.dir[2])));
        } else {
            auto al = std::make_shared<DiffuseAreaLight>(Transform(), MediumInterface(),
                                                         Spectrum::FromRGB(rgb), IILE_LIGHT_NSAMPLES(l),
                                                         shapes[l.prim], IILE_LIGHT_TWOSIDED(l));
            areaOf[l.prim] = al;
            rs.lights.push_back(al);
        }
    }
    for (size_t k = 0; k < fs.prims.size(); ++k) {
        std::shared_ptr<Material> m = fs.prims[k].material >= 0 ? mats[fs.prims[k].material] : nullptr;
        rs.prims[k] = std::make_shared<GeometricPrimitive>(shapes[k], m, areaOf[k], MediumInterface());
        rs.primIndex[rs.prims[k].get()] = (int)k;
    }
    BVHAccel::SplitMethod sm = BVHAccel::SplitMethod::SAH;
    if (fs.setup.splitMethod == IILE_SPLIT_MIDDLE) sm = BVHAccel::SplitMethod::Middle;
    if (fs.setup.splitMethod == IILE_SPLIT_EQUAL) sm = BVHAccel::SplitMethod::EqualCounts;
    if (fs.setup.splitMethod == IILE_SPLIT_HLBVH) sm = BVHAccel::SplitMethod::HLBVH;
    rs.bvh = std::make_shared<BVHAccel>(rs.prims, fs.setup.maxNodePrims, sm);
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- bvh --------------------------------------------------------------------
static int cmdBvh(int argc, char **argv) {
    if (argc < 5) die("usage: bvh <scene.flat> <nodes.bin> <order.bin>");
    Flat fs = loadFlat(argv[2]);
    bool needsInit = fs.setup.splitMethod == IILE_SPLIT_HLBVH;
    for (const IileLight &l : fs.lights) needsInit = needsInit || l.type == IILE_LIGHT_INFINITE;  // ctor uses ParallelFor
    if (needsInit) {
        // HLBVHBuild goes through ParallelFor, which needs pbrtInit; one thread = no workers, so the
        // atomic counter that hands leaves their slots in orderedPrims is walked in treelet order
        Options opt;
        opt.nThreads = 1;
        opt.quiet = true;
        pbrtInit(opt);
    }
    RefScene rs;
    double t0 = now();
    buildRefScene(fs, rs, false);
    double t1 = now();
    // count nodes by walking the flattened tree
    const LinearBVHNode *nodes = rs.bvh->nodes;
    size_t nNodes = 0;
    {
        std::vector<int> st{0};
        while (!st.empty()) {
            int c = st.back();
            st.pop_back();
            nNodes = std::max(nNodes, (size_t)c + 1);
            if (nodes[c].nPrimitives == 0) {
                st.push_back(c + 1);
                st.push_back(nodes[c].secondChildOffset);
            }
        }
    }
    std::ofstream fn(argv[3], std::ios::binary);
    fn.write((const char *)nodes, nNodes * sizeof(LinearBVHNode));
    std::vector<uint32_t> order;
    for (auto &p : rs.bvh->primitives) order.push_back((uint32_t)rs.primIndex[p.get()]);
    std::ofstream fo(argv[4], std::ios::binary);
    fo.write((const char *)order.data(), order.size() * 4);
    printf("{\"nodes\": %zu, \"prims\": %zu, \"build_plus_setup_s\": %.4f}\n", nNodes, order.size(), t1 - t0);
    return 0;
}

// ---- genrays ------------------------------------------------------------------
static int cmdGenRays(int argc, char **argv) {
    if (argc < 5) die("usage: genrays <scene.flat> <nrays> <rays.bin>");
    Flat fs = loadFlat(argv[2]);
    size_t n = (size_t)atoll(argv[3]);
    // world bound of the triangles only, expanded by half its diagonal
    Bounds3f b;
    for (const IilePrim &p : fs.prims)
        if (p.type == IILE_PRIM_TRIANGLE)
            for (int c = 0; c < 3; ++c)
                b = Union(b, Point3f(fs.P[3 * p.v[c]], fs.P[3 * p.v[c] + 1], fs.P[3 * p.v[c] + 2]));
    Vector3f half = b.Diagonal() * 0.5f;
    Point3f lo = b.pMin - half, hi = b.pMax + half;
    std::vector<IileRay> rays(n);
    for (size_t c = 0; c * 16384 < n; ++c) {
        RNG rng(c);
        for (size_t j = 0; j < 16384 && c * 16384 + j < n; ++j) {
            Float ox = rng.UniformFloat(), oy = rng.UniformFloat(), oz = rng.UniformFloat();
            Float u0 = rng.UniformFloat(), u1 = rng.UniformFloat();
            Vector3f d = UniformSampleSphere(Point2f(u0, u1));
            IileRay &r = rays[c * 16384 + j];
            r.o[0] = Lerp(ox, lo.x, hi.x);
            r.o[1] = Lerp(oy, lo.y, hi.y);
            r.o[2] = Lerp(oz, lo.z, hi.z);
            r.tMax = Infinity;
            r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z;
            r.time = 0;
        }
    }
    std::ofstream f(argv[4], std::ios::binary);
    f.write((const char *)rays.data(), rays.size() * sizeof(IileRay));
    printf("{\"rays\": %zu, \"lo\": [%.9g, %.9g, %.9g], \"hi\": [%.9g, %.9g, %.9g]}\n", n, lo.x, lo.y,
           lo.z, hi.x, hi.y, hi.z);
    return 0;
}

// ---- raycast ------------------------------------------------------------------
struct RefHit {  // 48 B
    int32_t prim;
    float t, b1, b2;
    float p[3];
    float n[3];
    float pad[2];
};

static int cmdRaycast(int argc, char **argv) {
    if (argc < 6) die("usage: raycast <scene.flat> <rays.bin> <out.bin> closest|any [nthreads]");
    Flat fs = loadFlat(argv[2]);
    bool any = std::string(argv[5]) == "any";
    int nThreads = argc > 6 ? atoi(argv[6]) : 0;
    Options opt;
    opt.nThreads = nThreads;
    opt.quiet = true;
    pbrtInit(opt);
    RefScene rs;
    buildRefScene(fs, rs, true);
    std::ifstream fr(argv[3], std::ios::binary | std::ios::ate);
    if (!fr) die(std::string("cannot open ") + argv[3]);
    size_t n = (size_t)fr.tellg() / sizeof(IileRay);
    fr.seekg(0);
    std::vector<IileRay> rays(n);
    fr.read((char *)rays.data(), n * sizeof(IileRay));
    std::vector<RefHit> hits(any ? 0 : n);
    std::vector<const Primitive *> hitPrim(any ? 0 : n, nullptr);  // mapped to indices after the timed loop
    std::vector<uint8_t> occ(any ? n : 0);
    const size_t chunk = 16384;
    size_t nChunks = (n + chunk - 1) / chunk;
    double t0 = now();
    ParallelFor(
        [&](int64_t c) {
            size_t end = std::min(n, (size_t)(c + 1) * chunk);
            for (size_t i = (size_t)c * chunk; i < end; ++i) {
                const IileRay &r = rays[i];
                Ray ray(Point3f(r.o[0], r.o[1], r.o[2]), Vector3f(r.d[0], r.d[1], r.d[2]), r.tMax, r.time);
                if (any) {
                    occ[i] = rs.bvh->IntersectP(ray) ? 1 : 0;
                } else {
                    SurfaceInteraction isect;
                    RefHit h;
                    std::memset(&h, 0, sizeof(h));
                    if (rs.bvh->Intersect(ray, &isect)) {
                        hitPrim[i] = isect.primitive;
                        h.prim = 0;
                        h.b1 = isect.uv.x;
                        h.b2 = isect.uv.y;
                        h.p[0] = isect.p.x; h.p[1] = isect.p.y; h.p[2] = isect.p.z;
                        h.n[0] = isect.n.x; h.n[1] = isect.n.y; h.n[2] = isect.n.z;
                    } else
                        h.prim = -1;
                    h.t = ray.tMax;
                    hits[i] = h;
                }
            }
        },
        (int64_t)nChunks);
    double t1 = now();
    for (size_t i = 0; i < hitPrim.size(); ++i)
        if (hitPrim[i]) hits[i].prim = rs.primIndex.at(hitPrim[i]);
    std::ofstream fo(argv[4], std::ios::binary);
    size_t nHit = 0;
    if (any) {
        fo.write((const char *)occ.data(), occ.size());
        for (uint8_t o : occ) nHit += o;
    } else {
        fo.write((const char *)hits.data(), hits.size() * sizeof(RefHit));
        for (const RefHit &h : hits) nHit += h.prim >= 0;
    }
    printf("{\"rays\": %zu, \"hits\": %zu, \"seconds\": %.6f, \"mrays_per_s\": %.4f, \"threads\": %d, \"kind\": \"%s\"}\n",
           n, nHit, t1 - t0, n / (t1 - t0) / 1e6, NumSystemCores(), any ? "any" : "closest");
    pbrtCleanup();
    return 0;
}

// ---- render -------------------------------------------------------------------
// One PCG32 stream per (pixel, sample): sequence index = sample << 32 | pixel, where
// pixel counts row-major inside Film::GetSampleBounds().  This is the sampler the GPU
// wavefront uses (it cannot consume one serial stream per 16x16 tile as RandomSampler
// does, random.cpp:56-60); installing it here makes both sides draw identical numbers.
class PixelStreamSampler : public Sampler {
  public:
    PixelStreamSampler(int64_t spp, const Bounds2i &sb, int sampleOffset)
        : Sampler(spp), sb(sb), sampleOffset(sampleOffset) {}
    void StartPixel(const Point2i &p) override {
        Sampler::StartPixel(p);
        fillArrays();
        reseed();
    }
    bool StartNextSample() override {
        bool more = Sampler::StartNextSample();
        if (more) reseed();
        return more;
    }
    Float Get1D() override { return rng.UniformFloat(); }
    Point2f Get2D() override {
        Float x = rng.UniformFloat();
        Float y = rng.UniformFloat();
        return Point2f(x, y);
    }
    // like RandomSampler::Clone (random.cpp:66-72: a copy of *this), a clone keeps the sample arrays the
    // integrator requested before Render() cloned the sampler (Request2DArray, sampler.cpp:75-80)
    std::unique_ptr<Sampler> Clone(int) override {
        PixelStreamSampler *c = new PixelStreamSampler(samplesPerPixel, sb, sampleOffset);
        c->samples1DArraySizes = samples1DArraySizes;
        c->samples2DArraySizes = samples2DArraySizes;
        c->sampleArray1D = sampleArray1D;
        c->sampleArray2D = sampleArray2D;
        return std::unique_ptr<Sampler>(c);
    }

  private:
    // The sample arrays of (pixel, sample) come from one more PCG32 stream, sequence = 1 << 62 | sample << 32 |
    // pixel, in the order the arrays were requested (RandomSampler::StartPixel fills them the same way from
    // its one serial stream, random.cpp:50-64): 1D arrays first, then the 2D arrays, x before y.
    void fillArrays() {
        if (sampleArray1D.empty() && sampleArray2D.empty()) return;
        uint64_t w = (uint64_t)(sb.pMax.x - sb.pMin.x);
        uint64_t pix = (uint64_t)(currentPixel.y - sb.pMin.y) * w + (uint64_t)(currentPixel.x - sb.pMin.x);
        for (int64_t s = 0; s < samplesPerPixel; ++s) {
            RNG a;
            a.SetSequence(((uint64_t)1 << 62) | ((uint64_t)(sampleOffset + s) << 32) | pix);
            for (size_t i = 0; i < sampleArray1D.size(); ++i)
                for (int k = 0; k < samples1DArraySizes[i]; ++k)
                    sampleArray1D[i][s * samples1DArraySizes[i] + k] = a.UniformFloat();
            for (size_t i = 0; i < sampleArray2D.size(); ++i)
                for (int k = 0; k < samples2DArraySizes[i]; ++k) {
                    Float x = a.UniformFloat();
                    Float y = a.UniformFloat();
                    sampleArray2D[i][s * samples2DArraySizes[i] + k] = Point2f(x, y);
                }
        }
    }
    void reseed() {
        uint64_t w = (uint64_t)(sb.pMax.x - sb.pMin.x);
        uint64_t pix = (uint64_t)(currentPixel.y - sb.pMin.y) * w + (uint64_t)(currentPixel.x - sb.pMin.x);
        uint64_t s = (uint64_t)(sampleOffset + currentPixelSampleIndex);
        rng.SetSequence((s << 32) | pix);
    }
    Bounds2i sb;
    int sampleOffset;
    RNG rng;
};

Sampler *makePixelStreamSampler(int64_t spp, const Bounds2i &sb, int sampleOffset) {
    return new PixelStreamSampler(spp, sb, sampleOffset);
}

static std::unique_ptr<Filter> makeFilter(const IileRenderSetup &s) {
    Vector2f r(s.filterRadius[0], s.filterRadius[1]);
    switch (s.filterType) {
    case IILE_FILTER_BOX: return std::unique_ptr<Filter>(new BoxFilter(r));
    case IILE_FILTER_GAUSSIAN: return std::unique_ptr<Filter>(new GaussianFilter(r, s.filterParam[0]));
    case IILE_FILTER_MITCHELL: return std::unique_ptr<Filter>(new MitchellFilter(r, s.filterParam[0], s.filterParam[1]));
    case IILE_FILTER_SINC: return std::unique_ptr<Filter>(new LanczosSincFilter(r, s.filterParam[0]));
    case IILE_FILTER_TRIANGLE: return std::unique_ptr<Filter>(new TriangleFilter(r));
    }
    die("unknown filter");
    return nullptr;
}

// the film and camera MakeFilm / MakeCamera build from the scene's setup (api.cpp:878-940)
std::shared_ptr<const Camera> makeCamera(const IileRenderSetup &S, const std::string &out, Film **filmOut) {
    Film *film = new Film(Point2i(S.xRes, S.yRes),
                          Bounds2f(Point2f(S.cropWindow[0], S.cropWindow[2]), Point2f(S.cropWindow[1], S.cropWindow[3])),
                          makeFilter(S), S.diagonal, out, S.filmScale, S.maxSampleLuminance);
    static std::vector<std::unique_ptr<Transform>> kept;  // AnimatedTransform keeps pointers
    kept.emplace_back(new Transform(mat(S.cameraToWorld), mat(S.worldToCamera)));
    const Transform *c2w = kept.back().get();
    AnimatedTransform ac2w(c2w, 0, c2w, 1);
    Bounds2f screen(Point2f(S.screenWindow[0], S.screenWindow[2]), Point2f(S.screenWindow[1], S.screenWindow[3]));
    std::shared_ptr<const Camera> camera;
    if (S.cameraType == IILE_CAMERA_ORTHOGRAPHIC)
        camera.reset(new OrthographicCamera(ac2w, screen, S.shutterOpen, S.shutterClose, S.lensRadius, S.focalDistance,
                                            film, nullptr));
    else if (S.cameraType == IILE_CAMERA_ENVIRONMENT)
        camera.reset(new EnvironmentCamera(ac2w, S.shutterOpen, S.shutterClose, film, nullptr));
    else
        camera.reset(new PerspectiveCamera(ac2w, screen, S.shutterOpen, S.shutterClose, S.lensRadius, S.focalDistance,
                                           S.fov, film, nullptr));
    *filmOut = film;
    return camera;
}

// REF_LOG_RAYS=<file>: every ray the integrator traces goes through this aggregate, which appends
// {o[3], tMax, d[3], kind (0 Intersect / 1 IntersectP), primitive hit (scene order, -1 none), t} to the file and
// forwards to the BVH (debugging aid for tools/fuzz_render.py's findings; one thread)
class LoggingAggregate : public Aggregate {
  public:
    LoggingAggregate(std::shared_ptr<Primitive> inner, const std::map<const Primitive *, int> *index, const char *path)
        : inner(inner), index(index), f(fopen(path, "wb")) {}
    ~LoggingAggregate() { if (f) fclose(f); }
    Bounds3f WorldBound() const override { return inner->WorldBound(); }
    bool Intersect(const Ray &r, SurfaceInteraction *isect) const override {
        float rec[10] = {r.o.x, r.o.y, r.o.z, r.tMax, r.d.x, r.d.y, r.d.z, 0.f, -1.f, 0.f};
        bool hit = inner->Intersect(r, isect);
        if (hit) {
            auto it = index->find(isect->primitive);
            rec[8] = it == index->end() ? -2.f : (float)it->second;
            rec[9] = r.tMax;
        }
        fwrite(rec, sizeof(rec), 1, f);
        return hit;
    }
    bool IntersectP(const Ray &r) const override {
        float rec[10] = {r.o.x, r.o.y, r.o.z, r.tMax, r.d.x, r.d.y, r.d.z, 1.f, -1.f, 0.f};
        bool hit = inner->IntersectP(r);
        rec[8] = hit ? 0.f : -1.f;
        fwrite(rec, sizeof(rec), 1, f);
        return hit;
    }

  private:
    std::shared_ptr<Primitive> inner;
    const std::map<const Primitive *, int> *index;
    FILE *f;
};

static int cmdRender(int argc, char **argv) {
    if (argc < 5) die("usage: render <scene.flat> <out.pfm> stock|pixelstream [nthreads] [spp] [sampleOffset] [repeats]");
    Flat fs = loadFlat(argv[2]);
    std::string out = argv[3];
    bool pixelStream = std::string(argv[4]) == "pixelstream";
    int nThreads = argc > 5 ? atoi(argv[5]) : 0;
    IileRenderSetup &S = fs.setup;
    if (argc > 6 && atoi(argv[6]) > 0) S.pixelSamples = atoi(argv[6]);
    int sampleOffset = argc > 7 ? atoi(argv[7]) : 0;
    // repeats > 1: Render() is called that many times on one scene (one BVH build for several timed
    // passes: bench.py's reference arm); the JSON line then carries every pass's seconds
    int repeats = argc > 8 ? std::max(1, atoi(argv[8])) : 1;
    Options opt;
    opt.nThreads = nThreads;
    opt.quiet = true;
    opt.imageFile = out;
    pbrtInit(opt);
    RefScene rs;
    double tb0 = now();
    buildRefScene(fs, rs, false);
    double tb1 = now();
    std::shared_ptr<Primitive> aggregate = rs.bvh;
    if (const char *logPath = getenv("REF_LOG_RAYS")) aggregate.reset(new LoggingAggregate(rs.bvh, &rs.primIndex, logPath));
    Scene scene(aggregate, rs.lights);

    Film *film = nullptr;
    std::shared_ptr<const Camera> camera = makeCamera(S, out, &film);
    Bounds2i sampleBounds = film->GetSampleBounds();
    std::shared_ptr<Sampler> sampler;
    if (pixelStream) sampler.reset(new PixelStreamSampler(S.pixelSamples, sampleBounds, sampleOffset));
    else if (S.samplerType == IILE_SAMPLER_HALTON)  // the reference's own HaltonSampler, as MakeSampler builds it
        sampler.reset(new HaltonSampler(S.pixelSamples, sampleBounds, S.samplePixelCenter != 0));
    else if (S.samplerType == IILE_SAMPLER_SOBOL)  // the reference's own SobolSampler (CreateSobolSampler, sobol.cpp:64-79)
        sampler.reset(new SobolSampler(S.pixelSamples, sampleBounds));
    else sampler.reset(new RandomSampler(S.pixelSamples));
    Bounds2i pixelBounds = sampleBounds;
    if (S.hasPixelBounds) {
        pixelBounds = Intersect(Bounds2i(Point2i(S.pixelBounds[0], S.pixelBounds[2]),
                                         Point2i(S.pixelBounds[1], S.pixelBounds[3])), sampleBounds);
    }
    const char *strategy = S.lightSampleStrategy == IILE_LIGHTSTRATEGY_UNIFORM ? "uniform"
                           : S.lightSampleStrategy == IILE_LIGHTSTRATEGY_POWER ? "power" : "spatial";
    std::unique_ptr<SamplerIntegrator> integratorPtr;
    if (S.integratorType == IILE_INTEGRATOR_DIRECT)
        integratorPtr.reset(new DirectLightingIntegrator(
            S.directStrategy == IILE_DIRECT_ONE ? LightStrategy::UniformSampleOne : LightStrategy::UniformSampleAll,
            S.maxDepth, camera, sampler, pixelBounds));
    else if (S.integratorType == IILE_INTEGRATOR_WHITTED)
        integratorPtr.reset(new WhittedIntegrator(S.maxDepth, camera, sampler, pixelBounds));
    else
        integratorPtr.reset(new PathIntegrator(S.maxDepth, camera, sampler, pixelBounds, S.rrThreshold, strategy));
    SamplerIntegrator &integrator = *integratorPtr;
    std::string passes;
    double t0 = 0, t1 = 0;
    for (int rep = 0; rep < repeats; ++rep) {
        t0 = now();
        integrator.Render(scene);
        t1 = now();
        char buf[64];
        snprintf(buf, sizeof(buf), "%s%.6f", rep ? ", " : "", t1 - t0);
        passes += buf;
    }
    MergeWorkerThreadStats();  // as pbrtWorldEnd does (api.cpp:1664-1667)
    ReportThreadStats();
    // the reference's own counters (stats.h), parsed by the caller
    FILE *sf = fopen((out + ".stats.txt").c_str(), "w");
    if (sf) {
        PrintStats(sf);
        fclose(sf);
    }
    int64_t nSamples = (int64_t)pixelBounds.Area() * S.pixelSamples;
    printf("{\"render_seconds\": %.6f, \"build_seconds\": %.4f, \"samples\": %lld, \"msamples_per_s\": %.5f, "
           "\"threads\": %d, \"sampler\": \"%s\", \"spp\": %d, \"pass_seconds\": [%s]}\n",
           t1 - t0, tb1 - tb0, (long long)nSamples, nSamples / (t1 - t0) / 1e6, NumSystemCores(),
           pixelStream ? "pixelstream" : (S.samplerType == IILE_SAMPLER_HALTON ? "stock-halton" : S.samplerType == IILE_SAMPLER_SOBOL ? "stock-sobol" : "stock"), S.pixelSamples,
           passes.c_str());
    pbrtCleanup();
    return 0;
}

int cmdHemi(int argc, char **argv);
int cmdHemiNN(int argc, char **argv);
int cmdIisptTask(int argc, char **argv);
int cmdMipmap(int argc, char **argv);
int cmdDirectPass(int argc, char **argv);
int cmdIileRender(int argc, char **argv);

int main(int argc, char **argv) {
    FLAGS_stderrthreshold = 2;
    if (argc < 2) die("usage: ref_harness bvh|raycast|render|genrays|hemi|hemi_nn|iispt_task|mipmap|direct_pass|iile_render ...");
    std::string cmd = argv[1];
    if (cmd == "bvh") return cmdBvh(argc, argv);
    if (cmd == "raycast") return cmdRaycast(argc, argv);
    if (cmd == "render") return cmdRender(argc, argv);
    if (cmd == "genrays") return cmdGenRays(argc, argv);
    if (cmd == "hemi") return cmdHemi(argc, argv);
    if (cmd == "hemi_nn") return cmdHemiNN(argc, argv);
    if (cmd == "iispt_task") return cmdIisptTask(argc, argv);
    if (cmd == "mipmap") return cmdMipmap(argc, argv);
    if (cmd == "direct_pass") return cmdDirectPass(argc, argv);
    if (cmd == "iile_render") return cmdIileRender(argc, argv);
    die("unknown command " + cmd);
    return 2;
}
