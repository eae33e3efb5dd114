// The opaque IileScene behind the C ABI: host-side flattened scene plus its
// device copy.
#pragma once

#include <mutex>
#include <string>
#include <vector>

#include "../host/film_host.h"
#include "../host/host_scene.h"
#include "device_scene.h"

struct IileScene {
    iile::HostScene host;
    iile::DeviceScene dev;
    std::vector<void *> deviceAllocs;
    bool uploaded = false;
    bool lightTablesReady = false;
    void *renderState = nullptr;  // wavefront buffers, owned by render_api.cu
    void *castStaging = nullptr;  // host-buffer cast pipeline (streams + staging), owned by c_api.cu
    // castPersistent's {work, done} counter pairs, one per CUDA stream that has launched a traversal
    // against this scene: launches on one stream are ordered and share a pair (the last block of a
    // launch puts it back to zero); launches on different streams may overlap and must not.
    // ... and, per stream as well, the bit set of rays a traversal launch leaves to the exact binary walk
    // (DeviceScene::deferBits), grown to the largest launch seen.
    struct StreamSlot {
        cudaStream_t stream;
        uint32_t *deferBits;
        size_t deferWords;
    };
    std::mutex counterMutex;
    std::vector<StreamSlot> counterStreams;
    ~IileScene();
};

namespace iile {
int setError(int code, const std::string &msg);
int cudaFail(cudaError_t e, const char *what);
int uploadScene(IileScene *s);
void IileSceneFree(IileScene *s);
void freeRenderState(IileScene *s);
void freeCastStaging(IileScene *s);
constexpr int kCounterPairs = 1024;
// The scene's device view with castCounters / deferBits pointing at `stream`'s own counter pair and bit
// set, the latter sized for launches of up to `maxItems` rays; IILE_E_UNSUPPORTED when more than
// kCounterPairs distinct streams have used the scene.
int deviceViewFor(IileScene *s, cudaStream_t stream, DeviceScene *out, size_t maxItems);
}  // namespace iile
