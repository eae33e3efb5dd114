// PCG32 generator, device side: the RNG behind RandomSampler
// (src/core/rng.h:62-156): 64-bit LCG state, xorshift-rotate output,
// SetSequence's two-step seeding, UniformFloat = min(1-eps, u32 * 2^-32), and
// the O(log n) jump-ahead of RNG::Advance.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace iile {

struct Pcg32 {
    uint64_t state, inc;
    static constexpr uint64_t kMult = 0x5851f42d4c957f2dULL;
    static constexpr uint64_t kDefaultState = 0x853c49e6748fea9bULL;

    __host__ __device__ void setSequence(uint64_t seq) {
        state = 0u;
        inc = (seq << 1u) | 1u;
        uniformUInt32();
        state += kDefaultState;
        uniformUInt32();
    }
    __host__ __device__ uint32_t uniformUInt32() {
        uint64_t old = state;
        state = old * kMult + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    __host__ __device__ float uniformFloat() {
        float f = (float)uniformUInt32() * 2.3283064365386963e-10f;  // 0x1p-32f
        const float oneMinusEps = 0.99999994f;                        // 0x1.fffffep-1
        return f < oneMinusEps ? f : oneMinusEps;
    }
    __host__ __device__ void advance(int64_t idelta) {
        uint64_t curMult = kMult, curPlus = inc, accMult = 1u, accPlus = 0u, delta = (uint64_t)idelta;
        while (delta > 0) {
            if (delta & 1) {
                accMult *= curMult;
                accPlus = accPlus * curMult + curPlus;
            }
            curPlus = (curMult + 1) * curPlus;
            curMult *= curMult;
            delta /= 2;
        }
        state = accMult * state + accPlus;
    }
};

}  // namespace iile
