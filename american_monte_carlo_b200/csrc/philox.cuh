// Counter-based Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) and the
// Box-Muller transforms used by the path generator.  __host__ __device__ so the CPU-only test tier can check
// the generator against the published known-answer vectors without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AMC_PHILOX_HD __host__ __device__ __forceinline__
#else
#define AMC_PHILOX_HD inline
#endif

namespace amc {

struct Philox4 {
    uint32_t v[4];
};

AMC_PHILOX_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

// ROUNDS = 10 is the generator of the paper and of cuRAND (the default everywhere in this library); the paper's
// Table 2 also lists Philox4x32-7 as the smallest round count that passes BigCrush -- offered as an explicit
// throughput option of the float path generator (AMC_PHILOX_ROUNDS=7), never used silently.
template <int ROUNDS>
AMC_PHILOX_HD Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int round = 0; round < ROUNDS; ++round) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c0, hi0, lo0);
        philox_mulhilo(M1, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 r;
    r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
    return r;
}

AMC_PHILOX_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}

// Counter layouts used by the path generators (documented in DESIGN.md); key = 64-bit user seed.
//   f64 storage:  c0, c1 = GLOBAL path id (low, high), c2 = block index along time (2 steps per call),
//                 c3 = 0x414d4331 ("AMC1")
//   f32 storage:  c0, c1 = GLOBAL path-QUAD id = path id / 4 (low, high), c2 = time step, c3 = 0x414d4332 ("AMC2"):
//                 one call yields the four normals of four ADJACENT paths at ONE step -- exactly what one thread
//                 stores as a 128-bit vector, and what the path-free backward sweep needs to regenerate step t
//                 without touching the other steps.
// Both depend on global ids only -> results do not depend on how paths are sharded over GPUs.
constexpr uint32_t kPhiloxDomain = 0x414d4331u;
constexpr uint32_t kPhiloxDomainQuad = 0x414d4332u;

}  // namespace amc
