#pragma once
#include <cuda_runtime.h>

#include "device_scene.h"

namespace iile {
cudaError_t launchClosestHit(const DeviceScene &sc, const IileRay *rays, size_t n, IileHit *hits,
                             cudaStream_t stream);
cudaError_t launchAnyHit(const DeviceScene &sc, const IileRay *rays, size_t n, uint8_t *occluded,
                         cudaStream_t stream);
cudaError_t launchGenerateRays(const float bmin[3], const float bmax[3], uint64_t firstRay, size_t n,
                               IileRay *rays, cudaStream_t stream);
}  // namespace iile
