// The opaque IileScene behind the C ABI: host-side flattened scene plus its
// device copy.
#pragma once

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
    ~IileScene();
};

namespace iile {
int setError(int code, const std::string &msg);
int cudaFail(cudaError_t e, const char *what);
int uploadScene(IileScene *s);
void IileSceneFree(IileScene *s);
void freeRenderState(IileScene *s);
void freeCastStaging(IileScene *s);
}  // namespace iile
