// Film::MergeFilmTile across GPUs (src/core/film.cpp:135-148): the per-rank {R,G,B,weight}
// films are summed with one NCCL collective over NVLink — SURVEY §8b.5, §8e.  The sum is the
// only exchange on the path (every (pixel, sample) is independent, the scene is replicated).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on
// it, a host that already carries NCCL — torch.distributed's bundled copy, or the system one —
// is reused by soname, and single-GPU users never load it.  Only the stable C entry points of
// nccl.h (2.x) are used; ncclUniqueId is the 128-byte blob the caller ships between ranks by
// whatever channel it has (file, socket, MPI, a torch.distributed broadcast).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "scene_handle.h"

using namespace iile;

namespace {

struct NcclId { char internal[IILE_COMM_ID_BYTES]; };  // ncclUniqueId, nccl.h: NCCL_UNIQUE_ID_BYTES = 128
typedef void *NcclComm;
// nccl.h enums (stable across 2.x): ncclFloat = 7, ncclSum = 0, ncclSuccess = 0
constexpr int kNcclFloat = 7, kNcclSum = 0;

struct NcclApi {
    void *handle = nullptr;
    int (*getVersion)(int *) = nullptr;
    int (*getUniqueId)(NcclId *) = nullptr;
    int (*commInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*commDestroy)(NcclComm) = nullptr;
    int (*commCount)(NcclComm, int *) = nullptr;
    int (*commUserRank)(NcclComm, int *) = nullptr;
    int (*allReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*reduce)(const void *, void *, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*getErrorString)(int) = nullptr;
    std::string loadError;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("IILE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
            api.loadError = dlerror();
        }
        if (!api.handle) return;
        auto sym = [&](const char *name) {
            void *p = dlsym(api.handle, name);
            if (!p) {
                api.loadError = std::string("missing symbol ") + name;
                api.handle = nullptr;
            }
            return p;
        };
#define BIND(field, name) \
    if (api.handle) *(void **)(&api.field) = sym(name)
        BIND(getVersion, "ncclGetVersion");
        BIND(getUniqueId, "ncclGetUniqueId");
        BIND(commInitRank, "ncclCommInitRank");
        BIND(commDestroy, "ncclCommDestroy");
        BIND(commCount, "ncclCommCount");
        BIND(commUserRank, "ncclCommUserRank");
        BIND(allReduce, "ncclAllReduce");
        BIND(reduce, "ncclReduce");
        BIND(getErrorString, "ncclGetErrorString");
#undef BIND
    });
    return api;
}

int needNccl(const char *who) {
    NcclApi &a = nccl();
    if (a.handle) return IILE_OK;
    return setError(IILE_E_UNSUPPORTED, std::string(who) + ": NCCL is not loadable (" + a.loadError +
                                            "); set IILE_NCCL_LIB to libnccl.so.2");
}

int ncclFail(int rc, const char *what) {
    NcclApi &a = nccl();
    return setError(IILE_E_CUDA, std::string(what) + ": NCCL error " + std::to_string(rc) + " (" +
                                     (a.getErrorString ? a.getErrorString(rc) : "?") + ")");
}

}  // namespace

extern "C" {

// ---- device buffers for hosts that do not link the CUDA runtime themselves ---------------------
int iile_dev_malloc(size_t bytes, void **ptr) {
    if (!ptr) return setError(IILE_E_INVALID, "iile_dev_malloc: null argument");
    *ptr = nullptr;
    cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 16);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, "iile_dev_malloc");
}
int iile_dev_free(void *ptr) {
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, "iile_dev_free");
}
int iile_dev_upload(void *dst, const void *src, size_t bytes) {
    if ((!dst || !src) && bytes) return setError(IILE_E_INVALID, "iile_dev_upload: null argument");
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, "iile_dev_upload");
}
int iile_dev_download(void *dst, const void *src, size_t bytes) {
    if ((!dst || !src) && bytes) return setError(IILE_E_INVALID, "iile_dev_download: null argument");
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, "iile_dev_download");
}
int iile_dev_synchronize(void *stream) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? IILE_OK : cudaFail(e, "iile_dev_synchronize");
}

int iile_comm_nccl_version(int *version) {
    int rc = needNccl("iile_comm_nccl_version");
    if (rc != IILE_OK) return rc;
    if (!version) return setError(IILE_E_INVALID, "iile_comm_nccl_version: null argument");
    const int e = nccl().getVersion(version);
    return e ? ncclFail(e, "ncclGetVersion") : IILE_OK;
}

int iile_comm_unique_id(uint8_t id[IILE_COMM_ID_BYTES]) {
    int rc = needNccl("iile_comm_unique_id");
    if (rc != IILE_OK) return rc;
    if (!id) return setError(IILE_E_INVALID, "iile_comm_unique_id: null argument");
    NcclId nid;
    const int e = nccl().getUniqueId(&nid);
    if (e) return ncclFail(e, "ncclGetUniqueId");
    memcpy(id, nid.internal, IILE_COMM_ID_BYTES);
    return IILE_OK;
}

int iile_comm_create(const uint8_t id[IILE_COMM_ID_BYTES], int nRanks, int rank, void **comm) {
    int rc = needNccl("iile_comm_create");
    if (rc != IILE_OK) return rc;
    if (!id || !comm) return setError(IILE_E_INVALID, "iile_comm_create: null argument");
    if (nRanks < 1 || rank < 0 || rank >= nRanks)
        return setError(IILE_E_INVALID, "iile_comm_create: rank outside [0, nRanks)");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev < 1)
        return setError(IILE_E_CUDA, "iile_comm_create: no CUDA device (there is no CPU path)");
    NcclId nid;
    memcpy(nid.internal, id, IILE_COMM_ID_BYTES);
    NcclComm c = nullptr;
    const int e = nccl().commInitRank(&c, nRanks, nid, rank);
    if (e) return ncclFail(e, "ncclCommInitRank");
    *comm = c;
    return IILE_OK;
}

int iile_comm_destroy(void *comm) {
    if (!comm) return IILE_OK;
    int rc = needNccl("iile_comm_destroy");
    if (rc != IILE_OK) return rc;
    const int e = nccl().commDestroy((NcclComm)comm);
    return e ? ncclFail(e, "ncclCommDestroy") : IILE_OK;
}

// Sum of the per-rank films, in place, asynchronous on `stream`: onto `root`, or onto every rank
// when root < 0.  A one-rank communicator (or comm == NULL) is the single-GPU case: nothing to do.
int iile_film_reduce_dev(float *filmRGBW, size_t nPixels, int root, void *comm, void *stream) {
    if (!filmRGBW && nPixels) return setError(IILE_E_INVALID, "iile_film_reduce_dev: null film");
    if (!comm || nPixels == 0) return IILE_OK;
    int rc = needNccl("iile_film_reduce_dev");
    if (rc != IILE_OK) return rc;
    NcclApi &a = nccl();
    int n = 0;
    int e = a.commCount((NcclComm)comm, &n);
    if (e) return ncclFail(e, "ncclCommCount");
    if (root >= n) return setError(IILE_E_INVALID, "iile_film_reduce_dev: root outside the communicator");
    if (n == 1) return IILE_OK;
    const size_t count = nPixels * 4;
    if (root < 0) e = a.allReduce(filmRGBW, filmRGBW, count, kNcclFloat, kNcclSum, (NcclComm)comm, (cudaStream_t)stream);
    else e = a.reduce(filmRGBW, filmRGBW, count, kNcclFloat, kNcclSum, root, (NcclComm)comm, (cudaStream_t)stream);
    return e ? ncclFail(e, root < 0 ? "ncclAllReduce" : "ncclReduce") : IILE_OK;
}
}
