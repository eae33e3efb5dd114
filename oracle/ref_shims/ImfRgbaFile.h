// Stand-in for OpenEXR's <ImfRgbaFile.h>; see ImfRgba.h in this directory.
#ifndef IILE_ORACLE_IMFRGBAFILE_SHIM_H
#define IILE_ORACLE_IMFRGBAFILE_SHIM_H
#include "ImfRgba.h"
#include <cstddef>
namespace Imf {
class RgbaInputFile {
  public:
    explicit RgbaInputFile(const char *) {
        throw std::runtime_error("EXR I/O is not available in the oracle/_ref build; use .pfm");
    }
    Imath::Box2i dataWindow() const { return Imath::Box2i(); }
    Imath::Box2i displayWindow() const { return Imath::Box2i(); }
    void setFrameBuffer(Rgba *, size_t, size_t) {}
    void readPixels(int, int) {}
};
class RgbaOutputFile {
  public:
    RgbaOutputFile(const char *, const Imath::Box2i &, const Imath::Box2i &, RgbaChannels) {
        throw std::runtime_error("EXR I/O is not available in the oracle/_ref build; use .pfm");
    }
    void setFrameBuffer(const Rgba *, size_t, size_t) {}
    void writePixels(int) {}
};
}  // namespace Imf
#endif
