// Shape "loopsubdiv": Loop subdivision to the limit surface, index-based.
//
// Same scheme, weights and evaluation order as the reference
// (src/shapes/loopsubdiv.cpp:149-400: even/odd vertex rules :226-275, topology
// update :279-321, limit positions :327-334, limit tangents/normals :336-372),
// but held in flat index arrays instead of a pointer graph: vertices and faces
// are integers, adjacency is `faceNbr[3*f+e]`, and the edge table is keyed by
// the (min,max) vertex-index pair — which equals the reference's pointer order
// for vertices allocated in index order.
#include <cmath>
#include <map>

#include "host_scene.h"

namespace iile {

namespace {

inline int nxt(int i) { return (i + 1) % 3; }
inline int prv(int i) { return (i + 2) % 3; }

struct Mesh {
    std::vector<V3> p;
    std::vector<int> startFace;           // per vertex
    std::vector<char> regular, boundary;  // per vertex
    std::vector<int> fv;                  // 3 per face: vertices
    std::vector<int> fn;                  // 3 per face: neighbour across edge (v[e], v[e+1]); -1 none

    int nFaces() const { return (int)fv.size() / 3; }
    int vnum(int f, int v) const {
        for (int i = 0; i < 3; ++i)
            if (fv[3 * f + i] == v) return i;
        throw std::runtime_error("loopsubdiv: inconsistent mesh topology");
    }
    int nextFace(int f, int v) const { return fn[3 * f + vnum(f, v)]; }
    int prevFace(int f, int v) const { return fn[3 * f + prv(vnum(f, v))]; }
    int nextVert(int f, int v) const { return fv[3 * f + nxt(vnum(f, v))]; }
    int prevVert(int f, int v) const { return fv[3 * f + prv(vnum(f, v))]; }
    int otherVert(int f, int a, int b) const {
        for (int i = 0; i < 3; ++i)
            if (fv[3 * f + i] != a && fv[3 * f + i] != b) return fv[3 * f + i];
        throw std::runtime_error("loopsubdiv: degenerate face");
    }
    int valence(int v) const {
        int f = startFace[v];
        if (!boundary[v]) {
            int n = 1;
            while ((f = nextFace(f, v)) != startFace[v]) ++n;
            return n;
        }
        int n = 1;
        while ((f = nextFace(f, v)) != -1) ++n;
        f = startFace[v];
        while ((f = prevFace(f, v)) != -1) ++n;
        return n + 1;
    }
    void oneRing(int v, std::vector<V3> &ring) const {
        ring.clear();
        if (!boundary[v]) {
            int f = startFace[v];
            do {
                ring.push_back(p[nextVert(f, v)]);
                f = nextFace(f, v);
            } while (f != startFace[v]);
        } else {
            int f = startFace[v], f2;
            while ((f2 = nextFace(f, v)) != -1) f = f2;
            ring.push_back(p[nextVert(f, v)]);
            do {
                ring.push_back(p[prevVert(f, v)]);
                f = prevFace(f, v);
            } while (f != -1);
        }
    }
};

inline float betaW(int valence) { return valence == 3 ? 3.f / 16.f : 3.f / (8.f * valence); }
inline float loopGamma(int valence) { return 1.f / (valence + 3.f / (8.f * betaW(valence))); }

V3 weightOneRing(const Mesh &m, int v, float beta, std::vector<V3> &ring) {
    int valence = m.valence(v);
    m.oneRing(v, ring);
    V3 p = (1 - valence * beta) * m.p[v];
    for (int i = 0; i < valence; ++i) p += beta * ring[i];
    return p;
}
V3 weightBoundary(const Mesh &m, int v, float beta, std::vector<V3> &ring) {
    int valence = m.valence(v);
    m.oneRing(v, ring);
    V3 p = (1 - 2 * beta) * m.p[v];
    p += beta * ring[0];
    p += beta * ring[valence - 1];
    return p;
}

using Edge = std::pair<int, int>;
inline Edge mkEdge(int a, int b) { return a < b ? Edge(a, b) : Edge(b, a); }

}  // namespace

void loopSubdivide(int nLevels, const std::vector<int> &indices, const std::vector<V3> &P,
                   std::vector<V3> *outP, std::vector<V3> *outN, std::vector<int> *outIdx) {
    Mesh m;
    int nVerts = (int)P.size();
    int nFaces = (int)indices.size() / 3;
    m.p = P;
    m.startFace.assign(nVerts, -1);
    m.regular.assign(nVerts, 0);
    m.boundary.assign(nVerts, 0);
    m.fv.assign(indices.begin(), indices.begin() + 3 * nFaces);
    m.fn.assign(3 * nFaces, -1);
    for (int f = 0; f < nFaces; ++f)
        for (int j = 0; j < 3; ++j) {
            int v = m.fv[3 * f + j];
            if (v < 0 || v >= nVerts) throw std::runtime_error("loopsubdiv: vertex index out of range");
            m.startFace[v] = f;
        }
    // neighbour pointers through an edge table (loopsubdiv.cpp:176-196)
    {
        std::map<Edge, std::pair<int, int>> open;  // edge -> (face, edgeNum)
        for (int f = 0; f < nFaces; ++f)
            for (int e = 0; e < 3; ++e) {
                Edge key = mkEdge(m.fv[3 * f + e], m.fv[3 * f + nxt(e)]);
                auto it = open.find(key);
                if (it == open.end())
                    open[key] = {f, e};
                else {
                    m.fn[3 * it->second.first + it->second.second] = f;
                    m.fn[3 * f + e] = it->second.first;
                    open.erase(it);
                }
            }
    }
    for (int v = 0; v < nVerts; ++v) {
        if (m.startFace[v] < 0) throw std::runtime_error("loopsubdiv: unreferenced control vertex");
        int f = m.startFace[v];
        do {
            f = m.nextFace(f, v);
        } while (f != -1 && f != m.startFace[v]);
        m.boundary[v] = (f == -1);
        int val = m.valence(v);
        m.regular[v] = (!m.boundary[v] && val == 6) || (m.boundary[v] && val == 4);
    }

    std::vector<V3> ring;
    for (int level = 0; level < nLevels; ++level) {
        Mesh c;  // next level
        int nv = (int)m.p.size(), nf = m.nFaces();
        // children of old vertices come first, in order; face f's children are 4f..4f+3
        c.p.resize(nv);
        c.regular.assign(m.regular.begin(), m.regular.end());
        c.boundary.assign(m.boundary.begin(), m.boundary.end());
        c.startFace.assign(nv, -1);
        c.fv.assign(12 * nf, -1);
        c.fn.assign(12 * nf, -1);
        // even vertices
        for (int v = 0; v < nv; ++v) {
            if (!m.boundary[v]) {
                if (m.regular[v]) c.p[v] = weightOneRing(m, v, 1.f / 16.f, ring);
                else c.p[v] = weightOneRing(m, v, betaW(m.valence(v)), ring);
            } else
                c.p[v] = weightBoundary(m, v, 1.f / 8.f, ring);
        }
        // odd (edge) vertices, in face/edge visiting order
        std::map<Edge, int> edgeVert;
        for (int f = 0; f < nf; ++f)
            for (int k = 0; k < 3; ++k) {
                int a = m.fv[3 * f + k], b = m.fv[3 * f + nxt(k)];
                Edge key = mkEdge(a, b);
                if (edgeVert.count(key)) continue;
                int id = (int)c.p.size();
                bool bnd = (m.fn[3 * f + k] == -1);
                V3 pos;
                int e0 = key.first, e1 = key.second;
                if (bnd) {
                    pos = 0.5f * m.p[e0];
                    pos += 0.5f * m.p[e1];
                } else {
                    pos = 3.f / 8.f * m.p[e0];
                    pos += 3.f / 8.f * m.p[e1];
                    pos += 1.f / 8.f * m.p[m.otherVert(f, e0, e1)];
                    pos += 1.f / 8.f * m.p[m.otherVert(m.fn[3 * f + k], e0, e1)];
                }
                c.p.push_back(pos);
                c.regular.push_back(1);
                c.boundary.push_back(bnd);
                c.startFace.push_back(4 * f + 3);
                edgeVert[key] = id;
            }
        // even vertex face pointers
        for (int v = 0; v < nv; ++v) {
            int sf = m.startFace[v];
            c.startFace[v] = 4 * sf + m.vnum(sf, v);
        }
        // face neighbour pointers
        for (int f = 0; f < nf; ++f)
            for (int j = 0; j < 3; ++j) {
                c.fn[3 * (4 * f + 3) + j] = 4 * f + nxt(j);
                c.fn[3 * (4 * f + j) + nxt(j)] = 4 * f + 3;
                int f2 = m.fn[3 * f + j];
                c.fn[3 * (4 * f + j) + j] = f2 != -1 ? 4 * f2 + m.vnum(f2, m.fv[3 * f + j]) : -1;
                f2 = m.fn[3 * f + prv(j)];
                c.fn[3 * (4 * f + j) + prv(j)] = f2 != -1 ? 4 * f2 + m.vnum(f2, m.fv[3 * f + j]) : -1;
            }
        // face vertex pointers
        for (int f = 0; f < nf; ++f)
            for (int j = 0; j < 3; ++j) {
                c.fv[3 * (4 * f + j) + j] = m.fv[3 * f + j];  // child of the even vertex keeps its id
                int ev = edgeVert[mkEdge(m.fv[3 * f + j], m.fv[3 * f + nxt(j)])];
                c.fv[3 * (4 * f + j) + nxt(j)] = ev;
                c.fv[3 * (4 * f + nxt(j)) + j] = ev;
                c.fv[3 * (4 * f + 3) + j] = ev;
            }
        m = std::move(c);
    }

    // limit surface positions
    int nv = (int)m.p.size();
    std::vector<V3> pLimit(nv);
    for (int v = 0; v < nv; ++v) {
        if (m.boundary[v]) pLimit[v] = weightBoundary(m, v, 1.f / 5.f, ring);
        else pLimit[v] = weightOneRing(m, v, loopGamma(m.valence(v)), ring);
    }
    m.p = pLimit;

    // limit tangents -> normals
    outN->clear();
    outN->reserve(nv);
    for (int v = 0; v < nv; ++v) {
        V3 S(0, 0, 0), T(0, 0, 0);
        int valence = m.valence(v);
        m.oneRing(v, ring);
        if (!m.boundary[v]) {
            for (int j = 0; j < valence; ++j) {
                S += std::cos(2 * kPi * j / valence) * ring[j];
                T += std::sin(2 * kPi * j / valence) * ring[j];
            }
        } else {
            S = ring[valence - 1] - ring[0];
            if (valence == 2)
                T = ring[0] + ring[1] - 2 * m.p[v];
            else if (valence == 3)
                T = ring[1] - m.p[v];
            else if (valence == 4)
                T = -1 * ring[0] + 2 * ring[1] + 2 * ring[2] + -1 * ring[3] + -2 * m.p[v];
            else {
                float theta = kPi / float(valence - 1);
                T = std::sin(theta) * (ring[0] + ring[valence - 1]);
                for (int k = 1; k < valence - 1; ++k) {
                    float wt = (2 * std::cos(theta) - 2) * std::sin((k)*theta);
                    T += wt * ring[k];
                }
                T = -T;
            }
        }
        outN->push_back(cross(S, T));
    }
    *outP = m.p;
    *outIdx = m.fv;
}

}  // namespace iile
