// HaltonSampler on the device (src/samplers/halton.cpp:63-122, GlobalSampler in
// src/core/sampler.cpp:128-196, radical inverses in src/core/lowdiscrepancy.cpp:388-425,
// 2490-2504, lowdiscrepancy.h:82-91).
//
// A Halton sample value depends only on (pixel, sample number, dimension), never on which
// thread or tile produced the previous one, so the GPU draws the very numbers the stock
// reference binary draws: dimension d of sample s of pixel p is the (scrambled) radical inverse
// in base prime(d) of   index = offsetForPixel(p) + s * sampleStride.
// PathIntegrator requests no sample arrays, so GlobalSampler::Get1D/Get2D simply walk the
// dimensions upwards (arrayStartDim == arrayEndDim == 5, sampler.cpp:180-194).
//
// Per path the stream is {index, next dimension}; it is kept in the 64-bit slot that holds the
// PCG32 state for the random sampler (index in bits 0..47, dimension in bits 48..63).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pcg.cuh"

namespace iile {

constexpr int kHaltonDims = 1000;       // PrimeTableSize, lowdiscrepancy.h:52
constexpr int kHaltonMaxResolution = 128;  // halton.cpp:42

// Device-resident tables: dimInfo[d] = {prime(d), PrimeSums[d]} and the digit permutations of
// ComputeRadicalInversePermutations(RNG()) (one shuffled 0..p-1 run per prime, back to back).
struct HaltonTables {
    uint2 dimInfo[kHaltonDims];
    const uint16_t *perms;
};

// SobolSampler (src/samplers/sobol.cpp:41-59, sobol.h:64-80; IILE's path integrator installs it when
// IILE_PATH_SAMPLES_OVERRIDE is set, path.cpp:201-212): also a GlobalSampler, so a value is a pure function
// of (pixel, sample number, dimension).  Tables: SobolMatrices32 [1024][52] and the rows m = log2Resolution
// of VdCSobolMatrices / VdCSobolMatricesInv (src/core/sobolmatrices.cpp; shipped as data/sobol_matrices.bin).
constexpr int kSobolDims = 1024;       // NumSobolDimensions, sobolmatrices.h:47
constexpr int kSobolMatrixSize = 52;   // SobolMatrixSize, sobolmatrices.h:48
struct SobolConsts {
    const uint32_t *matrices;  // nullptr => not the Sobol' sampler
    const uint64_t *vdc;       // [2][52]: VdCSobolMatrices[m - 1], VdCSobolMatricesInv[m - 1]
    int log2Resolution, resolution;
};

// The per-render constants of the HaltonSampler ctor (halton.cpp:63-91).
struct HaltonConsts {
    const HaltonTables *tables;  // nullptr => the random or the Sobol' sampler is in use
    SobolConsts sobol;
    int baseScales[2], baseExponents[2];
    int sampleStride;
    int multInverse[2];
    int samplePixelCenter;
};

__device__ __forceinline__ uint64_t inverseRadicalInverse(uint32_t base, uint64_t inverse, int nDigits) {
    uint64_t index = 0;
    for (int i = 0; i < nDigits; ++i) {
        const uint64_t digit = inverse % base;
        inverse /= base;
        index = index * base + digit;
    }
    return index;
}

// HaltonSampler::GetIndexForSample (halton.cpp:94-113): first index whose first two dimensions
// land in pixel (px, py) modulo kMaxResolution, plus sampleNum strides.
__device__ __forceinline__ uint64_t haltonIndexForSample(const HaltonConsts &h, int px, int py, uint32_t sampleNum) {
    uint64_t offset = 0;
    if (h.sampleStride > 1) {
        int pm[2] = {px % kHaltonMaxResolution, py % kHaltonMaxResolution};
        if (pm[0] < 0) pm[0] += kHaltonMaxResolution;  // Mod(), pbrt.h:311-314
        if (pm[1] < 0) pm[1] += kHaltonMaxResolution;
        const uint64_t d0 = inverseRadicalInverse(2, (uint64_t)pm[0], h.baseExponents[0]);
        const uint64_t d1 = inverseRadicalInverse(3, (uint64_t)pm[1], h.baseExponents[1]);
        offset += d0 * (uint64_t)(h.sampleStride / h.baseScales[0]) * (uint64_t)h.multInverse[0];
        offset += d1 * (uint64_t)(h.sampleStride / h.baseScales[1]) * (uint64_t)h.multInverse[1];
        offset %= (uint64_t)h.sampleStride;
    }
    return offset + (uint64_t)sampleNum * (uint64_t)h.sampleStride;
}

// RadicalInverse(0, a): `ReverseBits64(a) * 0x1p-64` is a double product returned as Float
// (lowdiscrepancy.cpp:429-435) — two roundings, both kept.
__device__ __forceinline__ float radicalInverseBase2(uint64_t a) {
    return (float)((double)__brevll(a) * 0x1p-64);
}

// RadicalInverseSpecialized<base> (lowdiscrepancy.cpp:388-403)
__device__ __forceinline__ float radicalInverse(uint32_t base, uint64_t a) {
    const float invBase = 1.f / (float)base;
    uint64_t reversed = 0;
    float invBaseN = 1;
    while (a) {
        const uint64_t next = a / base;
        const uint64_t digit = a - next * base;
        reversed = reversed * base + digit;
        invBaseN *= invBase;
        a = next;
    }
    return fminf((float)reversed * invBaseN, 0.99999994f);
}

// ScrambledRadicalInverseSpecialized<prime(dim)> (lowdiscrepancy.cpp:405-425) with
// HaltonSampler::PermutationForDimension (halton.h:71-76).  The reference aborts past
// dimension 999 (iile_render refuses such setups); the clamp only keeps the load in bounds.
static __device__ __noinline__ float haltonScrambled(const HaltonTables *__restrict__ t, uint64_t a, int dim) {
    const uint2 info = __ldg(&t->dimInfo[dim < kHaltonDims ? dim : kHaltonDims - 1]);
    const uint32_t base = info.x;
    const uint16_t *__restrict__ perm = t->perms + info.y;
    const float invBase = 1.f / (float)base;
    uint64_t reversed = 0;
    float invBaseN = 1;
    while (a >> 32) {
        const uint64_t next = a / base;
        const uint32_t digit = (uint32_t)(a - next * base);
        reversed = reversed * base + __ldg(perm + digit);
        invBaseN *= invBase;
        a = next;
    }
    uint32_t a32 = (uint32_t)a;
    while (a32) {
        const uint32_t next = a32 / base;
        const uint32_t digit = a32 - next * base;
        reversed = reversed * base + __ldg(perm + digit);
        invBaseN *= invBase;
        a32 = next;
    }
    const float tail = invBase * (float)__ldg(perm) / (1.f - invBase);
    return fminf(invBaseN * ((float)reversed + tail), 0.99999994f);
}

// SobolIntervalToIndex (lowdiscrepancy.h:229-249): index of sample `frame` of pixel (px, py) counted from
// the sample bounds' corner, for a 2^m x 2^m grid.
__device__ __forceinline__ uint64_t sobolIntervalToIndex(const SobolConsts &sb, uint64_t frame, uint32_t px, uint32_t py) {
    const uint32_t m = (uint32_t)sb.log2Resolution;
    if (m == 0) return 0;
    const uint32_t m2 = m << 1;
    uint64_t index = frame << m2;
    uint64_t delta = 0;
    for (int c = 0; frame; frame >>= 1, ++c)
        if (frame & 1) delta ^= __ldg(sb.vdc + c);
    uint64_t b = (((uint64_t)px << m) | (uint64_t)py) ^ delta;
    for (int c = 0; b; b >>= 1, ++c)
        if (b & 1) index ^= __ldg(sb.vdc + kSobolMatrixSize + c);
    return index;
}

// SobolSampleFloat (lowdiscrepancy.h:259-272), scramble 0
static __device__ __noinline__ float sobolSample(const uint32_t *__restrict__ matrices, uint64_t a, int dim) {
    uint32_t v = 0;
    const uint32_t *__restrict__ col = matrices + (dim < kSobolDims ? dim : kSobolDims - 1) * kSobolMatrixSize;
    for (; a != 0; a >>= 1, ++col)
        if (a & 1) v ^= __ldg(col);
    return fminf((float)v * 2.3283064365386963e-10f, 0.99999994f);
}

// What Li draws from: RNG::UniformFloat of the path's PCG32 stream (random sampler) or the next
// dimension of the path's Halton / Sobol' point.
struct SampleStream {
    Pcg32 pcg;                   // halton / sobol: pcg.state holds the sample index
    int dim;
    const HaltonTables *tables;  // Sampler "halton"
    const uint32_t *sobol;       // Sampler "sobol"; both nullptr => random

    __device__ __forceinline__ float uniformFloat() {
        if (tables) return haltonScrambled(tables, pcg.state, dim++);
        if (sobol) return sobolSample(sobol, pcg.state, dim++);
        return pcg.uniformFloat();
    }
    // the 64-bit word kept in PathState::rng between kernels
    __device__ __forceinline__ uint64_t pack() const {
        return (tables || sobol) ? (pcg.state | ((uint64_t)dim << 48)) : pcg.state;
    }
    __device__ __forceinline__ void unpack(uint64_t word, uint32_t sample, uint32_t pixel, const HaltonConsts &h) {
        tables = h.tables;
        sobol = h.sobol.matrices;
        if (tables || sobol) {
            pcg.state = word & 0xffffffffffffULL;
            pcg.inc = 0;
            dim = (int)(word >> 48);
        } else {
            pcg.state = word;
            pcg.inc = ((((uint64_t)sample << 32) | pixel) << 1u) | 1u;
            dim = 0;
        }
    }
};

}  // namespace iile
