// Stand-in for OpenEXR's <ImfRgba.h>, used ONLY so that the unmodified
// reference src/core/imageio.cpp compiles into oracle/_ref/ without the
// vendored OpenEXR cmake sub-project.  EXR I/O raises at run time; the oracle
// reads and writes PFM (the reference's own WriteImagePFM/ReadImagePFM).
#ifndef IILE_ORACLE_IMFRGBA_SHIM_H
#define IILE_ORACLE_IMFRGBA_SHIM_H
#include <stdexcept>
namespace Imath {
struct V2i {
    int x, y;
    V2i() : x(0), y(0) {}
    V2i(int x, int y) : x(x), y(y) {}
};
struct Box2i {
    V2i min, max;
    Box2i() {}
    Box2i(const V2i &a, const V2i &b) : min(a), max(b) {}
};
}  // namespace Imath
namespace Imf {
struct Rgba {
    float r, g, b, a;
    Rgba() : r(0), g(0), b(0), a(1) {}
    Rgba(float r, float g, float b, float a = 1.f) : r(r), g(g), b(b), a(a) {}
};
enum RgbaChannels { WRITE_RGBA = 0xf };
}  // namespace Imf
#endif
