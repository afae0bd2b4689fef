// The float path generator's per-step arithmetic, shared by the forward path kernel (pathgen.cu), the path-free
// backward sweep (lsm_sweep.cuh) and the generator self-test (amc_selftest_normals).
//
// One Philox4x32 call -> two Box-Muller pairs -> the log2-price increments of four adjacent paths at one step, as
// INTEGERS in units of 2^-k (fixed point).  A path's log2-price is the int32 running sum L_t of its increments:
//   * integer addition is exact and associative: no compensated summation, no drift over 252 steps;
//   * it is exactly reversible: L_{t-1} = L_t - q_t with q_t recomputed from the counter -- the path-free sweep walks the
//     SAME prices backwards bit for bit, so stored and regenerated paths take identical exercise decisions;
//   * S_t = S0 * 2^(L_t * 2^-k) is a pure function of L_t.
// k is chosen per path set (fixed_point_bits below) so that the largest possible increment stays below 2^22 (float ->
// int by the magic-number add, no conversion instruction) and |L| cannot overflow in any realistic excursion.  The
// quantisation of an increment is unbiased (round to nearest even) with standard deviation 2^-k / sqrt(12) in log2
// units: at config 3 (k = 25) 1.4e-7 after 252 steps, the size of the float rounding of the stored price itself.
// Everything stays on the FP32 / integer / MUFU pipes.
#pragma once
#include <math.h>
#include <stdint.h>

#include "philox.cuh"

namespace amc {

struct QuadGen {
    float kr;        // (-2 ln 2) * (vol_log2 * 2^k)^2: radius' = sqrt(kr * lg2(u1)) = vol_log2 * 2^k * sqrt(-2 ln u1)
    float dk;        // drift_log2 * 2^k
    float inv;       // 2^-k
    float S0;
    uint32_t k0, k1; // Philox key (the 64-bit seed)
    int k;           // fixed-point bits
};

constexpr float kBoxMullerMaxRadius = 6.8f;   // sqrt(-2 ln 2^-33) = 6.76: u1 >= 2^-33

// largest k such that (a) |increment| * 2^k < 2^22 and (b) an 8-sigma excursion of the whole path fits int32
inline int fixed_point_bits(double drift_log2, double vol_log2, int n_steps) {
    const double gmax = fabs(drift_log2) + vol_log2 * (double)kBoxMullerMaxRadius;
    int ka = 30;
    if (gmax > 0.0) ka = 22 - (int)ceil(log2(gmax));                    // gmax * 2^ka <= 2^22 (radius < 6.8: strict)
    const double span = (double)n_steps * fabs(drift_log2) + vol_log2 * (8.0 * sqrt((double)n_steps) + 7.0) + 1.0;
    const int kb = 30 - (int)ceil(log2(span));                          // span * 2^kb <= 2^30
    int k = ka < kb ? ka : kb;
    if (k > 30) k = 30;
    if (k < 4) k = 4;
    return k;
}

#if defined(__CUDACC__)
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr float kRoundMagic = 12582912.0f;          // 1.5 * 2^23: x + magic has round-to-nearest(x) in its mantissa
constexpr int kRoundMagicBits = 0x4B400000;

// standard normals of one Box-Muller pair, scaled by `scale` and shifted by `shift` (FMA-folded)
__device__ __forceinline__ void box_muller_pair(uint32_t a, uint32_t b, float kr, float shift, float& x, float& y) {
    // u1 in (0, 1]: full 32-bit resolution in the tail (small integers convert exactly)
    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float rad = sqrt_approx(kr * lg2_approx(u1));
    // angle: the low 23 bits become the mantissa of a float in [1, 2) (one LOP3, no int->float conversion);
    // 2 pi f - 3 pi lies in [-pi, pi)
    const float f12 = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
    const float ang = fmaf(f12, 6.283185307179586f, -9.42477796076938f);
    x = fmaf(rad, cos_approx(ang), shift);
    y = fmaf(rad, sin_approx(ang), shift);
}

// q[i] = fixed-point log2-price increment of path 4*quad + i at step t (t = 1..n: the step INTO column t)
template <int ROUNDS>
__device__ __forceinline__ void quad_increments(const QuadGen& g, uint32_t quad_lo, uint32_t quad_hi, uint32_t t,
                                                int (&q)[4]) {
    const Philox4 r = philox4x32<ROUNDS>(quad_lo, quad_hi, t, kPhiloxDomainQuad, g.k0, g.k1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float x, y;
        box_muller_pair(r.v[2 * h], r.v[2 * h + 1], g.kr, g.dk, x, y);
        q[2 * h] = __float_as_int(x + kRoundMagic) - kRoundMagicBits;
        q[2 * h + 1] = __float_as_int(y + kRoundMagic) - kRoundMagicBits;
    }
}

__device__ __forceinline__ float price_from_log(const QuadGen& g, int L) {
    return g.S0 * ex2_approx((float)L * g.inv);
}
#endif

}  // namespace amc
