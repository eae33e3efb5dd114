// Host-side float32 vector / matrix substrate for the scene front-end.
//
// Mirrors the arithmetic (operation order, float rounding points) of the
// reference's src/core/geometry.h and src/core/transform.{h,cpp} for the
// operations the flattening path needs: CTM composition, point / vector /
// normal transforms, LookAt / Rotate / Perspective, Gauss-Jordan inverse
// (transform.cpp:83-148), bounds unions.  All of this runs once per scene on
// the host; the results (world-space vertices, camera matrices) are what the
// kernels consume.  Build with -ffp-contract=off so that no a*b+c is fused.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>

namespace iile {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInfinity = std::numeric_limits<float>::infinity();
constexpr float kMachineEpsilon = std::numeric_limits<float>::epsilon() * 0.5f;
inline constexpr float gammaErr(int n) { return (n * kMachineEpsilon) / (1 - n * kMachineEpsilon); }
inline float radians(float deg) { return (kPi / 180.f) * deg; }

struct V3 {
    float x = 0, y = 0, z = 0;
    V3() {}
    V3(float x, float y, float z) : x(x), y(y), z(z) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(float s, const V3 &a) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(const V3 &a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
inline V3 &operator+=(V3 &a, const V3 &b) {
    a.x += b.x; a.y += b.y; a.z += b.z;
    return a;
}
inline float dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float lengthSquared(const V3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float length(const V3 &a) { return std::sqrt(lengthSquared(a)); }
// geometry.h:388-397 divides through a reciprocal (v * (1/len)).
inline V3 divide(const V3 &a, float f) {
    float inv = (float)1 / f;
    return V3(a.x * inv, a.y * inv, a.z * inv);
}
inline V3 normalize(const V3 &a) { return divide(a, length(a)); }
// geometry.h:957-963: the cross product is evaluated in double, rounded once.
inline V3 cross(const V3 &a, const V3 &b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V3((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)),
              (float)((ax * by) - (ay * bx)));
}

struct Bounds3 {
    V3 pMin{kInfinity, kInfinity, kInfinity};
    V3 pMax{-kInfinity, -kInfinity, -kInfinity};
    Bounds3() {}
    explicit Bounds3(const V3 &p) : pMin(p), pMax(p) {}
    Bounds3(const V3 &a, const V3 &b)
        : pMin(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)),
          pMax(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)) {}
    V3 diagonal() const { return pMax - pMin; }
    float surfaceArea() const {
        V3 d = diagonal();
        return 2 * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    int maximumExtent() const {
        V3 d = diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        return d.y > d.z ? 1 : 2;
    }
    // geometry.h Bounds3::Offset
    V3 offset(const V3 &p) const {
        V3 o = p - pMin;
        if (pMax.x > pMin.x) o.x /= pMax.x - pMin.x;
        if (pMax.y > pMin.y) o.y /= pMax.y - pMin.y;
        if (pMax.z > pMin.z) o.z /= pMax.z - pMin.z;
        return o;
    }
};
inline Bounds3 unite(const Bounds3 &b, const V3 &p) {
    Bounds3 r;
    r.pMin = V3(std::min(b.pMin.x, p.x), std::min(b.pMin.y, p.y), std::min(b.pMin.z, p.z));
    r.pMax = V3(std::max(b.pMax.x, p.x), std::max(b.pMax.y, p.y), std::max(b.pMax.z, p.z));
    return r;
}
inline Bounds3 unite(const Bounds3 &a, const Bounds3 &b) {
    Bounds3 r;
    r.pMin = V3(std::min(a.pMin.x, b.pMin.x), std::min(a.pMin.y, b.pMin.y),
                std::min(a.pMin.z, b.pMin.z));
    r.pMax = V3(std::max(a.pMax.x, b.pMax.x), std::max(a.pMax.y, b.pMax.y),
                std::max(a.pMax.z, b.pMax.z));
    return r;
}

struct M44 {
    float m[4][4];
    M44() {
        std::memset(m, 0, sizeof(m));
        m[0][0] = m[1][1] = m[2][2] = m[3][3] = 1.f;
    }
    M44(float a00, float a01, float a02, float a03, float a10, float a11, float a12, float a13,
        float a20, float a21, float a22, float a23, float a30, float a31, float a32, float a33) {
        float v[16] = {a00, a01, a02, a03, a10, a11, a12, a13,
                       a20, a21, a22, a23, a30, a31, a32, a33};
        std::memcpy(m, v, sizeof(m));
    }
    bool operator==(const M44 &o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
};

inline M44 mul(const M44 &a, const M44 &b) {
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] +
                        a.m[i][3] * b.m[3][j];
    return r;
}
inline M44 transpose(const M44 &a) {
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}

// Gauss-Jordan elimination with full pivoting, same pivot rule and update
// order as transform.cpp:83-148 so that inverses agree to the bit.
inline M44 inverse(const M44 &src) {
    int colIdx[4], rowIdx[4];
    int used[4] = {0, 0, 0, 0};
    float a[4][4];
    std::memcpy(a, src.m, sizeof(a));
    for (int step = 0; step < 4; ++step) {
        int pr = 0, pc = 0;
        float best = 0.f;
        for (int r = 0; r < 4; ++r) {
            if (used[r] == 1) continue;
            for (int c = 0; c < 4; ++c) {
                if (used[c] == 0) {
                    if (std::abs(a[r][c]) >= best) {
                        best = std::abs(a[r][c]);
                        pr = r;
                        pc = c;
                    }
                } else if (used[c] > 1)
                    throw std::runtime_error("Singular matrix in MatrixInvert");
            }
        }
        ++used[pc];
        if (pr != pc)
            for (int k = 0; k < 4; ++k) std::swap(a[pr][k], a[pc][k]);
        rowIdx[step] = pr;
        colIdx[step] = pc;
        if (a[pc][pc] == 0.f) throw std::runtime_error("Singular matrix in MatrixInvert");
        float pivInv = (float)(1. / a[pc][pc]);
        a[pc][pc] = 1.f;
        for (int k = 0; k < 4; ++k) a[pc][k] *= pivInv;
        for (int r = 0; r < 4; ++r) {
            if (r == pc) continue;
            float keep = a[r][pc];
            a[r][pc] = 0;
            for (int k = 0; k < 4; ++k) a[r][k] -= a[pc][k] * keep;
        }
    }
    for (int j = 3; j >= 0; --j)
        if (rowIdx[j] != colIdx[j])
            for (int k = 0; k < 4; ++k) std::swap(a[k][rowIdx[j]], a[k][colIdx[j]]);
    M44 out;
    std::memcpy(out.m, a, sizeof(a));
    return out;
}

// A matrix with its inverse carried alongside, as transform.h:126-226 does.
struct Xform {
    M44 m, inv;
    Xform() {}
    explicit Xform(const M44 &mm) : m(mm), inv(inverse(mm)) {}
    Xform(const M44 &mm, const M44 &ii) : m(mm), inv(ii) {}
    Xform operator*(const Xform &t) const { return Xform(mul(m, t.m), mul(t.inv, inv)); }
    bool operator==(const Xform &o) const { return m == o.m && inv == o.inv; }
    bool swapsHandedness() const {
        float det = m.m[0][0] * (m.m[1][1] * m.m[2][2] - m.m[1][2] * m.m[2][1]) -
                    m.m[0][1] * (m.m[1][0] * m.m[2][2] - m.m[1][2] * m.m[2][0]) +
                    m.m[0][2] * (m.m[1][0] * m.m[2][1] - m.m[1][1] * m.m[2][0]);
        return det < 0;
    }
    V3 point(const V3 &p) const {
        float x = p.x, y = p.y, z = p.z;
        float xp = m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z + m.m[0][3];
        float yp = m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z + m.m[1][3];
        float zp = m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z + m.m[2][3];
        float wp = m.m[3][0] * x + m.m[3][1] * y + m.m[3][2] * z + m.m[3][3];
        if (wp == 1) return V3(xp, yp, zp);
        return divide(V3(xp, yp, zp), wp);
    }
    V3 vector(const V3 &v) const {
        float x = v.x, y = v.y, z = v.z;
        return V3(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z,
                  m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z,
                  m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z);
    }
    V3 normal(const V3 &n) const {  // inverse transpose, transform.h:243-249
        float x = n.x, y = n.y, z = n.z;
        return V3(inv.m[0][0] * x + inv.m[1][0] * y + inv.m[2][0] * z,
                  inv.m[0][1] * x + inv.m[1][1] * y + inv.m[2][1] * z,
                  inv.m[0][2] * x + inv.m[1][2] * y + inv.m[2][2] * z);
    }
    Bounds3 bounds(const Bounds3 &b) const {  // transform.cpp:230-241 corner order
        Bounds3 r(point(V3(b.pMin.x, b.pMin.y, b.pMin.z)));
        r = unite(r, point(V3(b.pMax.x, b.pMin.y, b.pMin.z)));
        r = unite(r, point(V3(b.pMin.x, b.pMax.y, b.pMin.z)));
        r = unite(r, point(V3(b.pMin.x, b.pMin.y, b.pMax.z)));
        r = unite(r, point(V3(b.pMin.x, b.pMax.y, b.pMax.z)));
        r = unite(r, point(V3(b.pMax.x, b.pMax.y, b.pMin.z)));
        r = unite(r, point(V3(b.pMax.x, b.pMin.y, b.pMax.z)));
        r = unite(r, point(V3(b.pMax.x, b.pMax.y, b.pMax.z)));
        return r;
    }
};
inline Xform inverse(const Xform &t) { return Xform(t.inv, t.m); }

inline Xform xfTranslate(const V3 &d) {
    return Xform(M44(1, 0, 0, d.x, 0, 1, 0, d.y, 0, 0, 1, d.z, 0, 0, 0, 1),
                 M44(1, 0, 0, -d.x, 0, 1, 0, -d.y, 0, 0, 1, -d.z, 0, 0, 0, 1));
}
inline Xform xfScale(float x, float y, float z) {
    return Xform(M44(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1),
                 M44(1 / x, 0, 0, 0, 0, 1 / y, 0, 0, 0, 0, 1 / z, 0, 0, 0, 0, 1));
}
inline Xform xfRotate(float thetaDeg, const V3 &axis) {  // transform.cpp:182-205
    V3 a = normalize(axis);
    float s = std::sin(radians(thetaDeg));
    float c = std::cos(radians(thetaDeg));
    M44 r;
    r.m[0][0] = a.x * a.x + (1 - a.x * a.x) * c;
    r.m[0][1] = a.x * a.y * (1 - c) - a.z * s;
    r.m[0][2] = a.x * a.z * (1 - c) + a.y * s;
    r.m[0][3] = 0;
    r.m[1][0] = a.x * a.y * (1 - c) + a.z * s;
    r.m[1][1] = a.y * a.y + (1 - a.y * a.y) * c;
    r.m[1][2] = a.y * a.z * (1 - c) - a.x * s;
    r.m[1][3] = 0;
    r.m[2][0] = a.x * a.z * (1 - c) - a.y * s;
    r.m[2][1] = a.y * a.z * (1 - c) + a.x * s;
    r.m[2][2] = a.z * a.z + (1 - a.z * a.z) * c;
    r.m[2][3] = 0;
    return Xform(r, transpose(r));
}
inline Xform xfLookAt(const V3 &pos, const V3 &look, const V3 &up) {  // transform.cpp:207-241
    M44 c2w;
    c2w.m[0][3] = pos.x;
    c2w.m[1][3] = pos.y;
    c2w.m[2][3] = pos.z;
    c2w.m[3][3] = 1;
    V3 dir = normalize(look - pos);
    if (length(cross(normalize(up), dir)) == 0)
        throw std::runtime_error("LookAt: \"up\" vector and viewing direction are collinear");
    V3 right = normalize(cross(normalize(up), dir));
    V3 newUp = cross(dir, right);
    c2w.m[0][0] = right.x; c2w.m[1][0] = right.y; c2w.m[2][0] = right.z; c2w.m[3][0] = 0;
    c2w.m[0][1] = newUp.x; c2w.m[1][1] = newUp.y; c2w.m[2][1] = newUp.z; c2w.m[3][1] = 0;
    c2w.m[0][2] = dir.x;   c2w.m[1][2] = dir.y;   c2w.m[2][2] = dir.z;   c2w.m[3][2] = 0;
    return Xform(inverse(c2w), c2w);
}
inline Xform xfPerspective(float fovDeg, float n, float f) {  // transform.cpp:303-311
    M44 persp(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, f / (f - n), -f * n / (f - n), 0, 0, 1, 0);
    float invTan = 1 / std::tan(radians(fovDeg) / 2);
    return xfScale(invTan, invTan, 1) * Xform(persp);
}

inline float nextFloatUp(float v) {
    if (std::isinf(v) && v > 0.f) return v;
    if (v == -0.f) v = 0.f;
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if (v >= 0) ++u; else --u;
    std::memcpy(&v, &u, 4);
    return v;
}
inline float nextFloatDown(float v) {
    if (std::isinf(v) && v < 0.f) return v;
    if (v == 0.f) v = -0.f;
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if (v > 0) --u; else ++u;
    std::memcpy(&v, &u, 4);
    return v;
}

}  // namespace iile
