// oracle/cast_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Scalar CPU restatement of the reference's ray-cast path, used by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker for
// the CUDA kernels, and as the source of the per-ray node / triangle-test
// counts behind the roofline's algorithmic bytes.  Nothing in
// pbrt-v3-iile_b200/ links, imports or calls this file.
//
// Pinning: this restatement is checked bit for bit (hit flag, primitive,
// t bits, barycentric bits) against the reference itself — oracle/_ref's
// BVHAccel::Intersect / IntersectP driven by oracle/ref_harness — in
// tests/test_oracle_pinned.py, and against the committed golden vectors in
// tests/golden/ that were generated from it.
//
// Follows, function by function:
//   RNG (PCG32)                      src/core/rng.h:62-156
//   UniformSampleSphere              src/core/sampling.cpp:98-103
//   Bounds3::IntersectP              src/core/geometry.h:1412-1438
//   Triangle::Intersect (hit test)   src/shapes/triangle.cpp:188-291
//   Sphere::Intersect  (hit test)    src/shapes/sphere.cpp:49-103, EFloat src/core/efloat.h:48-283,
//                                    Transform(ray, &oErr, &dErr) src/core/transform.h:287-381
//   BVHAccel::Intersect / IntersectP src/accelerators/bvh.cpp:662-738
// Build: g++ -O2 -ffp-contract=off (no FMA contraction, like the reference).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr float kEps = std::numeric_limits<float>::epsilon() * 0.5f;
inline constexpr float gam(int n) { return (n * kEps) / (1 - n * kEps); }

struct Node {  // LinearBVHNode, bvh.cpp:95-104
    float bmin[3], bmax[3];
    int32_t offset;
    uint16_t nPrims;
    uint8_t axis, pad;
};
static_assert(sizeof(Node) == 32, "node size");

struct Ray {
    float o[3], tMax, d[3], time;
};
struct Hit {
    int32_t prim;
    float t, b1, b2;
};
struct SphereRec {  // IileSphere, include/iile_b200.h
    float o2w[16], w2o[16];
    float radius, zMin, zMax, thetaMin, thetaMax, phiMax;
    uint32_t flags;
    float phiMaxDeg;
};

// ---- PCG32 ------------------------------------------------------------------
struct Rng {
    uint64_t state, inc;
    explicit Rng(uint64_t seq) {
        state = 0u;
        inc = (seq << 1u) | 1u;
        next();
        state += 0x853c49e6748fea9bULL;
        next();
    }
    uint32_t next() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    float uniform() {
        float f = next() * 0x1p-32f;
        return f < 0x1.fffffep-1f ? f : 0x1.fffffep-1f;
    }
};

// ---- box ----------------------------------------------------------------------
inline bool slab(const Node &n, const float o[3], const float inv[3], const int neg[3], float tMaxRay) {
    const float *b[2] = {n.bmin, n.bmax};
    float tMin = (b[neg[0]][0] - o[0]) * inv[0];
    float tMax = (b[1 - neg[0]][0] - o[0]) * inv[0];
    float tyMin = (b[neg[1]][1] - o[1]) * inv[1];
    float tyMax = (b[1 - neg[1]][1] - o[1]) * inv[1];
    tMax *= 1 + 2 * gam(3);
    tyMax *= 1 + 2 * gam(3);
    if (tMin > tyMax || tyMin > tMax) return false;
    if (tyMin > tMin) tMin = tyMin;
    if (tyMax < tMax) tMax = tyMax;
    float tzMin = (b[neg[2]][2] - o[2]) * inv[2];
    float tzMax = (b[1 - neg[2]][2] - o[2]) * inv[2];
    tzMax *= 1 + 2 * gam(3);
    if (tMin > tzMax || tzMin > tMax) return false;
    if (tzMin > tMin) tMin = tzMin;
    if (tzMax < tMax) tMax = tzMax;
    return (tMin < tMaxRay) && (tMax > 0);
}

// ---- triangle -------------------------------------------------------------------
inline float max3(float a, float b, float c) { return std::max(a, std::max(b, c)); }

bool triangle(const float o[3], const float d[3], float tMax, const float *p0, const float *p1,
              const float *p2, float *tHit, float b[3]) {
    float p0t[3] = {p0[0] - o[0], p0[1] - o[1], p0[2] - o[2]};
    float p1t[3] = {p1[0] - o[0], p1[1] - o[1], p1[2] - o[2]};
    float p2t[3] = {p2[0] - o[0], p2[1] - o[1], p2[2] - o[2]};
    float ad[3] = {std::abs(d[0]), std::abs(d[1]), std::abs(d[2])};
    int kz = (ad[0] > ad[1]) ? ((ad[0] > ad[2]) ? 0 : 2) : ((ad[1] > ad[2]) ? 1 : 2);
    int kx = kz + 1;
    if (kx == 3) kx = 0;
    int ky = kx + 1;
    if (ky == 3) ky = 0;
    float dp[3] = {d[kx], d[ky], d[kz]};
    float a0[3] = {p0t[kx], p0t[ky], p0t[kz]};
    float a1[3] = {p1t[kx], p1t[ky], p1t[kz]};
    float a2[3] = {p2t[kx], p2t[ky], p2t[kz]};
    float Sx = -dp[0] / dp[2], Sy = -dp[1] / dp[2], Sz = 1.f / dp[2];
    a0[0] += Sx * a0[2];
    a0[1] += Sy * a0[2];
    a1[0] += Sx * a1[2];
    a1[1] += Sy * a1[2];
    a2[0] += Sx * a2[2];
    a2[1] += Sy * a2[2];
    float e0 = a1[0] * a2[1] - a1[1] * a2[0];
    float e1 = a2[0] * a0[1] - a2[1] * a0[0];
    float e2 = a0[0] * a1[1] - a0[1] * a1[0];
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        double p2txp1ty = (double)a2[0] * (double)a1[1];
        double p2typ1tx = (double)a2[1] * (double)a1[0];
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)a0[0] * (double)a2[1];
        double p0typ2tx = (double)a0[1] * (double)a2[0];
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)a1[0] * (double)a0[1];
        double p1typ0tx = (double)a1[1] * (double)a0[0];
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
    float det = e0 + e1 + e2;
    if (det == 0) return false;
    a0[2] *= Sz;
    a1[2] *= Sz;
    a2[2] *= Sz;
    float tScaled = e0 * a0[2] + e1 * a1[2] + e2 * a2[2];
    if (det < 0 && (tScaled >= 0 || tScaled < tMax * det)) return false;
    else if (det > 0 && (tScaled <= 0 || tScaled > tMax * det)) return false;
    float invDet = 1 / det;
    float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
    float t = tScaled * invDet;
    float maxZt = max3(std::abs(a0[2]), std::abs(a1[2]), std::abs(a2[2]));
    float deltaZ = gam(3) * maxZt;
    float maxXt = max3(std::abs(a0[0]), std::abs(a1[0]), std::abs(a2[0]));
    float maxYt = max3(std::abs(a0[1]), std::abs(a1[1]), std::abs(a2[1]));
    float deltaX = gam(5) * (maxXt + maxZt);
    float deltaY = gam(5) * (maxYt + maxZt);
    float deltaE = 2 * (gam(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
    float maxE = max3(std::abs(e0), std::abs(e1), std::abs(e2));
    float deltaT = 3 * (gam(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * std::abs(invDet);
    if (t <= deltaT) return false;
    *tHit = t;
    b[0] = b0;
    b[1] = b1;
    b[2] = b2;
    return true;
}

// ---- EFloat + sphere --------------------------------------------------------------
inline float up(float v) {
    if (std::isinf(v) && v > 0.f) return v;
    if (v == -0.f) v = 0.f;
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if (v >= 0) ++u; else --u;
    std::memcpy(&v, &u, 4);
    return v;
}
inline float down(float v) {
    if (std::isinf(v) && v < 0.f) return v;
    if (v == 0.f) v = -0.f;
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if (v > 0) --u; else ++u;
    std::memcpy(&v, &u, 4);
    return v;
}
struct EF {
    float v, lo, hi;
    EF() : v(0), lo(0), hi(0) {}
    EF(float v_, float err = 0.f) : v(v_) {
        if (err == 0.) lo = hi = v_;
        else {
            lo = down(v_ - err);
            hi = up(v_ + err);
        }
    }
    EF operator+(EF e) const {
        EF r;
        r.v = v + e.v;
        r.lo = down(lo + e.lo);
        r.hi = up(hi + e.hi);
        return r;
    }
    EF operator-(EF e) const {
        EF r;
        r.v = v - e.v;
        r.lo = down(lo - e.hi);
        r.hi = up(hi - e.lo);
        return r;
    }
    EF operator*(EF e) const {
        EF r;
        r.v = v * e.v;
        float p[4] = {lo * e.lo, hi * e.lo, lo * e.hi, hi * e.hi};
        r.lo = down(std::min(std::min(p[0], p[1]), std::min(p[2], p[3])));
        r.hi = up(std::max(std::max(p[0], p[1]), std::max(p[2], p[3])));
        return r;
    }
    EF operator/(EF e) const {
        EF r;
        r.v = v / e.v;
        if (e.lo < 0 && e.hi > 0) {
            r.lo = -std::numeric_limits<float>::infinity();
            r.hi = std::numeric_limits<float>::infinity();
        } else {
            float q[4] = {lo / e.lo, hi / e.lo, lo / e.hi, hi / e.hi};
            r.lo = down(std::min(std::min(q[0], q[1]), std::min(q[2], q[3])));
            r.hi = up(std::max(std::max(q[0], q[1]), std::max(q[2], q[3])));
        }
        return r;
    }
};

bool efQuadratic(EF A, EF B, EF C, EF *t0, EF *t1) {
    double discrim = (double)B.v * (double)B.v - 4. * (double)A.v * (double)C.v;
    if (discrim < 0.) return false;
    double root = std::sqrt(discrim);
    EF fr((float)root, (float)(kEps * root));
    EF q;
    if (B.v < 0) q = EF(-.5f) * (B - fr);
    else q = EF(-.5f) * (B + fr);
    *t0 = q / A;
    *t1 = C / q;
    if (t0->v > t1->v) std::swap(*t0, *t1);
    return true;
}

bool sphere(const SphereRec &s, const float ro[3], const float rd[3], float tMax, float *tHit,
            float uv[2]) {
    const float *m = s.w2o;
    const float g3 = gam(3);
    float x = ro[0], y = ro[1], z = ro[2];
    float ox = m[0] * x + m[1] * y + m[2] * z + m[3];
    float oy = m[4] * x + m[5] * y + m[6] * z + m[7];
    float oz = m[8] * x + m[9] * y + m[10] * z + m[11];
    float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    float oe[3] = {g3 * (std::abs(m[0] * x) + std::abs(m[1] * y) + std::abs(m[2] * z) + std::abs(m[3])),
                   g3 * (std::abs(m[4] * x) + std::abs(m[5] * y) + std::abs(m[6] * z) + std::abs(m[7])),
                   g3 * (std::abs(m[8] * x) + std::abs(m[9] * y) + std::abs(m[10] * z) + std::abs(m[11]))};
    if (wp != 1) {
        float inv = 1.f / wp;
        ox *= inv;
        oy *= inv;
        oz *= inv;
    }
    x = rd[0];
    y = rd[1];
    z = rd[2];
    float de[3] = {g3 * (std::abs(m[0] * x) + std::abs(m[1] * y) + std::abs(m[2] * z)),
                   g3 * (std::abs(m[4] * x) + std::abs(m[5] * y) + std::abs(m[6] * z)),
                   g3 * (std::abs(m[8] * x) + std::abs(m[9] * y) + std::abs(m[10] * z))};
    float dx = m[0] * x + m[1] * y + m[2] * z;
    float dy = m[4] * x + m[5] * y + m[6] * z;
    float dz = m[8] * x + m[9] * y + m[10] * z;
    float len2 = dx * dx + dy * dy + dz * dz;
    if (len2 > 0) {
        float dt = (std::abs(dx) * oe[0] + std::abs(dy) * oe[1] + std::abs(dz) * oe[2]) / len2;
        ox += dx * dt;
        oy += dy * dt;
        oz += dz * dt;
    }
    if (s.flags & 4u) {  // Shape "disk": Disk::Intersect, disk.cpp:48-80 (height in zMin, inner radius in zMax)
        const float height = s.zMin, inner = s.zMax;
        if (dz == 0) return false;
        float t = (height - oz) / dz;
        if (t <= 0 || t >= tMax) return false;
        float hx = ox + dx * t, hy = oy + dy * t;
        float dist2 = hx * hx + hy * hy;
        if (dist2 > s.radius * s.radius || dist2 < inner * inner) return false;
        float phi = std::atan2(hy, hx);
        if (phi < 0) phi += 2 * 3.14159265358979323846f;
        if (phi > s.phiMax) return false;
        *tHit = t;
        float rHit = std::sqrt(dist2);
        uv[0] = phi / s.phiMax;
        uv[1] = 1 - ((rHit - inner) / (s.radius - inner));
        return true;
    }
    if (s.flags & 8u) {  // Shape "cylinder": Cylinder::Intersect, cylinder.cpp:48-106
        EF eox(ox, oe[0]), eoy(oy, oe[1]);
        EF edx(dx, de[0]), edy(dy, de[1]);
        EF a = edx * edx + edy * edy;
        EF b = EF(2.f) * (edx * eox + edy * eoy);
        EF c = eox * eox + eoy * eoy - EF(s.radius) * EF(s.radius);
        EF t0, t1;
        if (!efQuadratic(a, b, c, &t0, &t1)) return false;
        if (t0.hi > tMax || t1.lo <= 0) return false;
        EF th = t0;
        if (th.lo <= 0) {
            th = t1;
            if (th.hi > tMax) return false;
        }
        float px, py, pz, phi;
        auto refine = [&](float t) {
            px = ox + dx * t;
            py = oy + dy * t;
            pz = oz + dz * t;
            float hitRad = std::sqrt(px * px + py * py);
            px *= s.radius / hitRad;
            py *= s.radius / hitRad;
            phi = std::atan2(py, px);
            if (phi < 0) phi += 2 * 3.14159265358979323846f;
        };
        refine(th.v);
        auto clipped = [&]() { return pz < s.zMin || pz > s.zMax || phi > s.phiMax; };
        if (clipped()) {
            if (th.v == t1.v) return false;
            th = t1;
            if (t1.hi > tMax) return false;
            refine(th.v);
            if (clipped()) return false;
        }
        *tHit = th.v;
        uv[0] = phi / s.phiMax;
        uv[1] = (pz - s.zMin) / (s.zMax - s.zMin);
        return true;
    }
    EF eox(ox, oe[0]), eoy(oy, oe[1]), eoz(oz, oe[2]);
    EF edx(dx, de[0]), edy(dy, de[1]), edz(dz, de[2]);
    EF a = edx * edx + edy * edy + edz * edz;
    EF b = EF(2.f) * (edx * eox + edy * eoy + edz * eoz);
    EF c = eox * eox + eoy * eoy + eoz * eoz - EF(s.radius) * EF(s.radius);
    EF t0, t1;
    if (!efQuadratic(a, b, c, &t0, &t1)) return false;
    if (t0.hi > tMax || t1.lo <= 0) return false;
    EF th = t0;
    if (th.lo <= 0) {
        th = t1;
        if (th.hi > tMax) return false;
    }
    float px, py, pz, phi;
    auto refine = [&](float t) {
        px = ox + dx * t;
        py = oy + dy * t;
        pz = oz + dz * t;
        float k = s.radius / std::sqrt(px * px + py * py + pz * pz);
        px *= k;
        py *= k;
        pz *= k;
        if (px == 0 && py == 0) px = 1e-5f * s.radius;
        phi = std::atan2(py, px);
        if (phi < 0) phi += 2 * 3.14159265358979323846f;
    };
    refine(th.v);
    auto clipped = [&]() {
        return (s.zMin > -s.radius && pz < s.zMin) || (s.zMax < s.radius && pz > s.zMax) || phi > s.phiMax;
    };
    if (clipped()) {
        if (th.v == t1.v) return false;
        if (t1.hi > tMax) return false;
        th = t1;
        refine(th.v);
        if (clipped()) return false;
    }
    *tHit = th.v;
    float cz = pz / s.radius;
    cz = cz < -1 ? -1 : (cz > 1 ? 1 : cz);
    uv[0] = phi / s.phiMax;
    uv[1] = (std::acos(cz) - s.thetaMin) / (s.thetaMax - s.thetaMin);
    return true;
}

struct SceneView {
    const Node *nodes;
    const float *P;            // vertex pool
    const int32_t *slotPrim;   // [nPrims] BVH slot -> scene primitive
    const int32_t *primType;   // [nPrims] by scene primitive
    const int32_t *primV;      // [nPrims][3]
    const SphereRec *spheres;
};

// bvh.cpp:662-738.  counts[0] += nodes fetched, counts[1] += primitive tests.
template <bool ANY>
bool traverse(const SceneView &sc, const Ray &ray, Hit *hit, uint64_t counts[2]) {
    float tMax = ray.tMax;
    float inv[3] = {1 / ray.d[0], 1 / ray.d[1], 1 / ray.d[2]};
    int neg[3] = {inv[0] < 0, inv[1] < 0, inv[2] < 0};
    int stack[64], sp = 0, cur = 0;
    bool found = false;
    if (hit) {
        hit->prim = -1;
        hit->b1 = hit->b2 = 0;
    }
    while (true) {
        const Node &n = sc.nodes[cur];
        ++counts[0];
        if (slab(n, ray.o, inv, neg, tMax)) {
            if (n.nPrims > 0) {
                for (int i = 0; i < n.nPrims; ++i) {
                    int prim = sc.slotPrim[n.offset + i];
                    ++counts[1];
                    float t, b[3] = {0, 0, 0}, uv[2];
                    bool h;
                    if (sc.primType[prim] == 0) {
                        const int32_t *v = sc.primV + 3 * prim;
                        h = triangle(ray.o, ray.d, tMax, sc.P + 3 * v[0], sc.P + 3 * v[1], sc.P + 3 * v[2], &t, b);
                    } else {
                        h = sphere(sc.spheres[sc.primV[3 * prim]], ray.o, ray.d, tMax, &t, uv);
                        b[1] = uv[0];
                        b[2] = uv[1];
                    }
                    if (h) {
                        if (ANY) return true;
                        found = true;
                        tMax = t;
                        hit->prim = prim;
                        hit->b1 = b[1];
                        hit->b2 = b[2];
                    }
                }
                if (sp == 0) break;
                cur = stack[--sp];
            } else {
                if (neg[n.axis]) {
                    stack[sp++] = cur + 1;
                    cur = n.offset;
                } else {
                    stack[sp++] = n.offset;
                    cur = cur + 1;
                }
            }
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    if (hit) hit->t = tMax;
    return found;
}

}  // namespace

extern "C" {

// cfg-2 ray set: chunk c of 2^14 rays uses RNG(c); five draws per ray.
void oracle_generate_rays(const float lo[3], const float hi[3], uint64_t first, uint64_t n, Ray *rays) {
    uint64_t i = 0;
    while (i < n) {
        uint64_t g = first + i;
        uint64_t c = g >> 14, j0 = g & 16383;
        Rng rng(c);
        for (uint64_t k = 0; k < 5 * j0; ++k) rng.next();
        for (uint64_t j = j0; j < 16384 && i < n; ++j, ++i) {
            float u[5];
            for (float &x : u) x = rng.uniform();
            Ray &r = rays[i];
            for (int a = 0; a < 3; ++a) r.o[a] = (1 - u[a]) * lo[a] + u[a] * hi[a];
            float z = 1 - 2 * u[3];
            float rr = std::sqrt(std::max(0.f, 1.f - z * z));
            float phi = 2 * 3.14159265358979323846f * u[4];
            r.d[0] = rr * std::cos(phi);
            r.d[1] = rr * std::sin(phi);
            r.d[2] = z;
            r.tMax = std::numeric_limits<float>::infinity();
            r.time = 0;
        }
    }
}

void oracle_pcg32_floats(uint64_t seq, uint64_t skip, uint64_t n, float *out) {
    Rng rng(seq);
    for (uint64_t k = 0; k < skip; ++k) rng.next();
    for (uint64_t k = 0; k < n; ++k) out[k] = rng.uniform();
}

// any = 0: hits[n] (prim, t = ray.tMax afterwards, b1, b2); any = 1: occluded[n].
// counts[2]: total nodes fetched and primitive tests.  nThreads <= 0: all cores.
void oracle_cast(const Node *nodes, const float *P, const int32_t *slotPrim, const int32_t *primType,
                 const int32_t *primV, const SphereRec *spheres, const Ray *rays, uint64_t n, int any,
                 Hit *hits, uint8_t *occluded, uint64_t counts[2], int nThreads) {
    SceneView sc{nodes, P, slotPrim, primType, primV, spheres};
    if (nThreads <= 0) nThreads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    std::vector<uint64_t> c(2 * (size_t)nThreads, 0);
    for (int t = 0; t < nThreads; ++t)
        pool.emplace_back([&, t]() {
            uint64_t lo = n * t / nThreads, hi = n * (t + 1) / nThreads;
            for (uint64_t i = lo; i < hi; ++i) {
                if (any) occluded[i] = traverse<true>(sc, rays[i], nullptr, &c[2 * t]) ? 1 : 0;
                else traverse<false>(sc, rays[i], &hits[i], &c[2 * t]);
            }
        });
    for (auto &th : pool) th.join();
    counts[0] = counts[1] = 0;
    for (int t = 0; t < nThreads; ++t) {
        counts[0] += c[2 * t];
        counts[1] += c[2 * t + 1];
    }
}

}  // extern "C"
