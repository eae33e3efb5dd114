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
    bool hasInterfaces = false;   // some primitive has no material (interface only): paths may take more steps than vertices
    int classesPresent = 0x3f;    // shading classes (wavefront.h kClass*) that occur in the scene, bit per class
    bool simpleLights = false;    // every light is a point light or a diffuse area light on a triangle
    void *renderState = nullptr;  // wavefront buffers, owned by render_api.cu
    void *castStaging = nullptr;  // host-buffer cast pipeline (streams + staging), owned by c_api.cu
    // castPersistent's {work, done} counter pairs, one per CUDA stream that has launched a traversal
    // against this scene: launches on one stream are ordered and share a pair (the last block of a
    // launch puts it back to zero); launches on different streams may overlap and must not.
    std::mutex counterMutex;
    std::vector<cudaStream_t> counterStreams;
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
// The scene's device view with castCounters pointing at `stream`'s own counter pair; IILE_E_UNSUPPORTED
// when more than kCounterPairs distinct streams have used the scene.
int deviceViewFor(IileScene *s, cudaStream_t stream, DeviceScene *out);
}  // namespace iile
