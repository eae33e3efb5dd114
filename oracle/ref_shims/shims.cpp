// Definitions backing oracle/ref_shims/glog/logging.h and the one reference
// translation unit left out of the oracle/_ref build (src/textures/ptex.cpp,
// which needs the vendored Ptex cmake sub-project).  Test infrastructure only.
#include <glog/logging.h>

#include "pbrt.h"
#include "error.h"
#include "paramset.h"
#include "textures/ptex.h"

int FLAGS_stderrthreshold = 2;
int FLAGS_minloglevel = 0;
int FLAGS_v = 0;
bool FLAGS_logtostderr = false;
std::string FLAGS_log_dir;

namespace pbrt {
PtexTexture<Float> *CreatePtexFloatTexture(const Transform &, const TextureParams &) {
    Error("ptex textures are not available in the oracle/_ref build");
    return nullptr;
}
PtexTexture<Spectrum> *CreatePtexSpectrumTexture(const Transform &, const TextureParams &) {
    Error("ptex textures are not available in the oracle/_ref build");
    return nullptr;
}
}  // namespace pbrt
