// Path simulation and path-matrix plumbing for sm_100a.
//
// generate_asset_paths (amc.py:72-81) draws Z[P, n], forms exp((r - sigma^2/2) dt + sigma sqrt(dt) Z) and
// takes a cumulative product along time into a path-major [P, n+1] f64 array (plus five same-size
// temporaries).  Here one kernel does all of it and writes the only layout the backward sweep wants:
// TIMESTEP-MAJOR S[t][p], so that every per-step access of the sweep is a contiguous column.
//   philox_paths_kernel  : K1  -- Philox4x32-10 counter = (global path id, time block), Box-Muller,
//                          cumulative sum in LOG space (double accumulator), one 128-bit store per thread and
//                          step (4 f32 paths or 2 f64 paths per thread).
//   normals_paths_kernel : K1z -- same arithmetic in f64 from caller-supplied normals Z[p][j] (row-major,
//                          what amc.py:74 draws), staged through shared memory to turn the path-major read
//                          into timestep-major coalesced writes.  The A/B mode against the reference.
//   transpose_in_kernel  : adopt a reference-layout path matrix S[p][t].
// plus column statistics, the knock-in index of the down-and-in barrier (amc.py:171-176) and read-back helpers.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "gbm_quad.cuh"
#include "philox.cuh"

namespace amc {

// ---------------------------------------------------------------------------------------------------------
// K1, f32 storage: one thread = one QUAD of four adjacent paths, walked through all n steps; one Philox call per quad
// and step (gbm_quad.cuh), the log2-prices kept as exact int32 fixed-point sums, one 128-bit streaming store per
// thread and step.  FP32 / integer / MUFU pipes only.  Per path-step: 10 (Philox-10) + 5 (Box-Muller) + 6 (sum, price)
// instructions, 3 MUFU ops.
//   S == nullptr: nothing is stored per step (path-free mode, amc_paths_generate_lean): only the terminal log-prices
//   L_n go to `L_out` -- the backward sweep regenerates every earlier column from them.
// MINB = 6: no register cap -- the natural allocation (34 registers, no spills) already gives 6 resident blocks per SM;
// MINB = 8 caps at 32 registers (a few spills) for 8 blocks.  The launcher sizes the grid as ONE resident wave of MINB
// blocks per SM.
template <int ROUNDS, bool ALIGNED, int MINB>
__global__ void __launch_bounds__(256, (MINB == 8 ? 8 : 1)) philox_quads_f32_kernel(float* __restrict__ S, int32_t* __restrict__ L_out,
                                                               int64_t ld, int n_steps, int64_t n_local,
                                                               int64_t path_offset, QuadGen g) {
    const int64_t quad0 = path_offset >> 2;                        // first global quad that holds a local path
    const int shift = (int)(path_offset & 3);                      // ALIGNED: 0
    const int64_t n_quads = ((path_offset + n_local + 3) >> 2) - quad0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_quads; v += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t quad = (uint64_t)(quad0 + v);
        const uint32_t qlo = (uint32_t)quad, qhi = (uint32_t)(quad >> 32);
        const int64_t p0 = (v << 2) - shift;                       // local index of the quad's first path (may be < 0)
        int L[4] = {0, 0, 0, 0};
        if (S) {
            if (ALIGNED) {
                st_stream(reinterpret_cast<float4*>(S + p0), make_float4(g.S0, g.S0, g.S0, g.S0));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (p0 + i >= 0 && p0 + i < n_local) S[p0 + i] = g.S0;
            }
        }
        float* out_col = S + p0;
        for (int t = 1; t <= n_steps; ++t) {
            int q[4];
            quad_increments<ROUNDS>(g, qlo, qhi, (uint32_t)t, q);
            float out[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                L[i] += q[i];
                out[i] = price_from_log(g, L[i]);
            }
            if (S) {
                out_col += ld;
                if (ALIGNED) {
                    st_stream(reinterpret_cast<float4*>(out_col), make_float4(out[0], out[1], out[2], out[3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (p0 + i >= 0 && p0 + i < n_local) out_col[i] = out[i];
                }
            }
        }
        if (L_out) {
            if (ALIGNED) {
                *reinterpret_cast<int4*>(L_out + p0) = make_int4(L[0], L[1], L[2], L[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (p0 + i >= 0 && p0 + i < n_local) L_out[p0 + i] = L[i];
            }
        }
    }
}

// K1, f64 storage: 2 adjacent paths per thread, 2 steps per Philox call and path (53-bit uniforms).
__global__ void __launch_bounds__(256) philox_paths_f64_kernel(double* __restrict__ S, int64_t ld, int n_steps,
                                                               int64_t n_local, int64_t path_offset, GbmParams g,
                                                               uint32_t k0, uint32_t k1) {
    const int64_t n_vec = (n_local + 1) >> 1;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p0 = v << 1;
        double L[2] = {0.0, 0.0};
        st_stream(reinterpret_cast<double2*>(S + p0), make_double2(g.S0, g.S0));
        for (int t0 = 0; t0 < n_steps; t0 += 2) {
            double z[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint64_t gid = (uint64_t)(path_offset + p0 + i);
                const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(t0 >> 1),
                                                kPhiloxDomain, k0, k1);
                const uint64_t a = ((uint64_t)r.v[0] << 32) | r.v[1];
                const uint64_t b = ((uint64_t)r.v[2] << 32) | r.v[3];
                const double u1 = ((double)(a >> 11) + 0.5) * 1.1102230246251565e-16;     // (0, 1)
                const double u2 = (double)(b >> 11) * 1.1102230246251565e-16;             // [0, 1)
                const double rad = sqrt(-2.0 * log(u1));
                double sn, cs;
                sincospi(2.0 * u2, &sn, &cs);
                z[i][0] = rad * cs;
                z[i][1] = rad * sn;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (t0 + j < n_steps) {
                    L[0] += fma(g.vol, z[0][j], g.drift);
                    L[1] += fma(g.vol, z[1][j], g.drift);
                    st_stream(reinterpret_cast<double2*>(S + (int64_t)(t0 + j + 1) * ld + p0),
                              make_double2(g.S0 * exp(L[0]), g.S0 * exp(L[1])));
                }
            }
        }
    }
}

QuadGen make_quad_gen(const GbmParams& g, int n_steps, uint64_t seed) {
    const double log2e = 1.4426950408889634;
    const double d2 = g.drift * log2e, v2 = fabs(g.vol) * log2e;     // log2 units; -z is as normal as z
    QuadGen q;
    q.k = fixed_point_bits(d2, v2, n_steps);
    const double sc = ldexp(1.0, q.k);
    q.kr = (float)(-1.3862943611198906 * (v2 * sc) * (v2 * sc));     // -2 ln u = (-2 ln 2) log2 u
    q.dk = (float)(d2 * sc);
    q.inv = (float)ldexp(1.0, -q.k);
    q.S0 = (float)g.S0;
    q.k0 = (uint32_t)seed;
    q.k1 = (uint32_t)(seed >> 32);
    return q;
}

int philox_rounds() {
    static const int rounds = [] {
        const char* e = getenv("AMC_PHILOX_ROUNDS");
        return (e && atoi(e) == 7) ? 7 : 10;
    }();
    return rounds;
}

// S == nullptr with L_out != nullptr: path-free mode (only the terminal log-prices are stored)
cudaError_t launch_generate_philox(int dtype, void* S, int32_t* L_out, int64_t ld, int n_steps, int64_t n_local,
                                   int64_t path_offset, GbmParams g, uint64_t seed, int sm_count, cudaStream_t s) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int64_t n_vec = dtype == 1 ? ((path_offset + n_local + 3) / 4 - path_offset / 4) : (n_local + 1) / 2;
    // ONE wave of resident blocks (a second, partial wave would run at a fraction of the occupancy for as long as the
    // first: ncu r2f_k1 showed 56 % warps active with the grid capped at 8 blocks per SM while only 6 were resident),
    // and the same number of grid-stride iterations for every thread
    static const int k1_blocks = (getenv("AMC_K1_BLOCKS") && atoi(getenv("AMC_K1_BLOCKS")) == 8) ? 8 : 6;
    const int per_sm = dtype == 1 ? (((path_offset & 3) == 0) ? k1_blocks : 6) : 4;              // __launch_bounds__(256, MINB) / the f64 kernel's registers
    const int64_t cap = (int64_t)sm_count * per_sm;
    int64_t blocks = (n_vec + 255) / 256;
    if (blocks > cap) {
        const int64_t iters = (n_vec + cap * 256 - 1) / (cap * 256);
        blocks = (n_vec + 256 * iters - 1) / (256 * iters);
    }
    if (blocks < 1) blocks = 1;
    const int b = (int)blocks;
    if (dtype == 1) {
        const QuadGen q = make_quad_gen(g, n_steps, seed);
        const bool aligned = (path_offset & 3) == 0;
        const bool seven = philox_rounds() == 7;
        float* Sf = (float*)S;
        if (aligned && !seven && k1_blocks == 8) philox_quads_f32_kernel<10, true, 8><<<b, 256, 0, s>>>(Sf, L_out, ld, n_steps, n_local, path_offset, q);
        else if (aligned && !seven) philox_quads_f32_kernel<10, true, 6><<<b, 256, 0, s>>>(Sf, L_out, ld, n_steps, n_local, path_offset, q);
        else if (aligned && k1_blocks == 8) philox_quads_f32_kernel<7, true, 8><<<b, 256, 0, s>>>(Sf, L_out, ld, n_steps, n_local, path_offset, q);
        else if (aligned) philox_quads_f32_kernel<7, true, 6><<<b, 256, 0, s>>>(Sf, L_out, ld, n_steps, n_local, path_offset, q);
        else if (!seven) philox_quads_f32_kernel<10, false, 6><<<b, 256, 0, s>>>(Sf, L_out, ld, n_steps, n_local, path_offset, q);
        else philox_quads_f32_kernel<7, false, 6><<<b, 256, 0, s>>>(Sf, L_out, ld, n_steps, n_local, path_offset, q);
    } else {
        if (L_out || !S) return cudaErrorInvalidValue;
        philox_paths_f64_kernel<<<b, 256, 0, s>>>((double*)S, ld, n_steps, n_local, path_offset, g, k0, k1);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Lean (path-free) sets keep no matrix: reading a column, rows, the knock-in index or each path's price at its own step
// walks the quads forward again from L_0 = 0 (diagnostics / API parity; the sweep itself walks backward from L_n).
enum LeanWalk { kLeanColumn = 0, kLeanRows = 1, kLeanFirstHit = 2, kLeanGather = 3 };

template <int ROUNDS, int MODE>
__global__ void __launch_bounds__(256) lean_walk_kernel(QuadGen g, int64_t quad0, int64_t n_local, int n_steps, int t_stop,
                                                        int64_t p_lo, int64_t p_hi, double barrier,
                                                        const int32_t* __restrict__ steps, double* __restrict__ out,
                                                        int32_t* __restrict__ out_i) {
    const int64_t q_lo = p_lo >> 2, q_hi = (p_hi + 3) >> 2;
    for (int64_t v = q_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < q_hi; v += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t quad = (uint64_t)(quad0 + v);
        int L[4] = {0, 0, 0, 0};
        int fh[4] = {n_steps + 1, n_steps + 1, n_steps + 1, n_steps + 1};
        int want[4] = {0, 0, 0, 0};
        double got[4] = {(double)g.S0, (double)g.S0, (double)g.S0, (double)g.S0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t pp = 4 * v + i;
            if (MODE == kLeanFirstHit && (double)g.S0 <= barrier) fh[i] = 0;
            if (MODE == kLeanGather && pp >= p_lo && pp < p_hi) {
                int w = steps[pp];
                want[i] = w < 0 ? 0 : (w > n_steps ? n_steps : w);
            }
            if (MODE == kLeanRows && pp >= p_lo && pp < p_hi) out[(pp - p_lo) * (int64_t)(n_steps + 1)] = (double)g.S0;
        }
        for (int t = 1; t <= t_stop; ++t) {
            int q[4];
            quad_increments<ROUNDS>(g, (uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)t, q);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                L[i] += q[i];
                const int64_t pp = 4 * v + i;
                if (MODE == kLeanColumn) continue;
                const double x = (double)price_from_log(g, L[i]);
                if (MODE == kLeanRows && pp >= p_lo && pp < p_hi) out[(pp - p_lo) * (int64_t)(n_steps + 1) + t] = x;
                if (MODE == kLeanFirstHit && x <= barrier && fh[i] == n_steps + 1) fh[i] = t;
                if (MODE == kLeanGather && t == want[i]) got[i] = x;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t pp = 4 * v + i;
            if (pp < p_lo || pp >= p_hi || pp >= n_local) continue;
            if (MODE == kLeanColumn) out[pp] = t_stop == 0 ? (double)g.S0 : (double)price_from_log(g, L[i]);
            if (MODE == kLeanFirstHit) out_i[pp] = fh[i];
            if (MODE == kLeanGather) out[pp] = got[i];
        }
    }
}

cudaError_t launch_lean_walk(int mode, int rounds, const QuadGen& g, int64_t quad0, int64_t n_local, int n_steps, int t_stop,
                             int64_t p_lo, int64_t p_hi, double barrier, const int32_t* steps_dev, double* out_dev,
                             int32_t* out_i_dev, int sm_count, cudaStream_t s) {
    const int64_t nq = ((p_hi + 3) >> 2) - (p_lo >> 2);
    int64_t blocks = (nq + 255) / 256;
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    if (blocks < 1) blocks = 1;
    const int b = (int)blocks;
#define AMC_LEAN_WALK(R, M) \
    lean_walk_kernel<R, M><<<b, 256, 0, s>>>(g, quad0, n_local, n_steps, t_stop, p_lo, p_hi, barrier, steps_dev, out_dev, out_i_dev)
    if (rounds == 7) {
        switch (mode) {
            case kLeanColumn: AMC_LEAN_WALK(7, kLeanColumn); break;
            case kLeanRows: AMC_LEAN_WALK(7, kLeanRows); break;
            case kLeanFirstHit: AMC_LEAN_WALK(7, kLeanFirstHit); break;
            case kLeanGather: AMC_LEAN_WALK(7, kLeanGather); break;
            default: return cudaErrorInvalidValue;
        }
    } else {
        switch (mode) {
            case kLeanColumn: AMC_LEAN_WALK(10, kLeanColumn); break;
            case kLeanRows: AMC_LEAN_WALK(10, kLeanRows); break;
            case kLeanFirstHit: AMC_LEAN_WALK(10, kLeanFirstHit); break;
            case kLeanGather: AMC_LEAN_WALK(10, kLeanGather); break;
            default: return cudaErrorInvalidValue;
        }
    }
#undef AMC_LEAN_WALK
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Generator self-tests (amc_selftest_*): the DEVICE build of the integer Philox stage for chosen counters (known-answer
// vectors must come out bit-exact -- philox.cuh takes a different mulhilo branch under __CUDA_ARCH__), and the
// distribution of the float Box-Muller normals exactly as the path kernel forms them (MUFU lg2 / sqrt / sin / cos).
template <int ROUNDS>
__global__ void philox_kat_kernel(const uint32_t* __restrict__ ctr, uint32_t k0, uint32_t k1, int n, uint32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox4 r = philox4x32<ROUNDS>(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], k0, k1);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[4 * i + j] = r.v[j];
}

cudaError_t launch_philox_kat(int rounds, const uint32_t* ctr_dev, uint32_t k0, uint32_t k1, int n, uint32_t* out_dev,
                              cudaStream_t s) {
    const int b = (n + 127) / 128;
    if (rounds == 10) philox_kat_kernel<10><<<b, 128, 0, s>>>(ctr_dev, k0, k1, n, out_dev);
    else if (rounds == 7) philox_kat_kernel<7><<<b, 128, 0, s>>>(ctr_dev, k0, k1, n, out_dev);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// standard normals of n_quads x n_steps Philox calls (4 normals each), binned on [lo, hi) into n_bins equal bins plus an
// underflow (bin 0) and an overflow (bin n_bins + 1) bin; stats[block][6] = count, sum z, z^2, z^3, z^4, max |z|
constexpr int kNormalBinsMax = 4096;
template <int ROUNDS>
__global__ void __launch_bounds__(256) normals_hist_kernel(uint32_t k0, uint32_t k1, int64_t n_quads, int n_steps, int n_bins,
                                                           float lo, float inv_width, unsigned long long* __restrict__ hist,
                                                           double* __restrict__ stats) {
    __shared__ unsigned int sh[kNormalBinsMax + 2];
    __shared__ double red[8 * 6];
    for (int i = threadIdx.x; i < n_bins + 2; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    float zmax = 0.f;
    int since_flush = 0;
    // block-uniform trip count (the flush below has block-wide barriers)
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n_quads; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = base + threadIdx.x;
        for (int t = 1; t <= n_steps && v < n_quads; ++t) {
            const Philox4 r = philox4x32<ROUNDS>((uint32_t)v, (uint32_t)((uint64_t)v >> 32), (uint32_t)t, kPhiloxDomainQuad, k0, k1);
            float z[4];
            box_muller_pair(r.v[0], r.v[1], -1.3862943611198906f, 0.f, z[0], z[1]);
            box_muller_pair(r.v[2], r.v[3], -1.3862943611198906f, 0.f, z[2], z[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float b = floorf((z[i] - lo) * inv_width);
                const int bin = b < 0.f ? 0 : (b >= (float)n_bins ? n_bins + 1 : (int)b + 1);
                atomicAdd(&sh[bin], 1u);
                const double zd = (double)z[i], z2 = zd * zd;
                acc[0] += 1.0; acc[1] += zd; acc[2] += z2; acc[3] += z2 * zd; acc[4] += z2 * z2;
                zmax = fmaxf(zmax, fabsf(z[i]));
            }
        }
        since_flush += n_steps;
        if (since_flush >= 2048) {                             // flush long before a 32-bit bin can overflow
            since_flush = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n_bins + 2; i += blockDim.x) {
                if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
                sh[i] = 0u;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins + 2; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
    double all[6] = {acc[0], acc[1], acc[2], acc[3], acc[4], 0.0};
    // max |z|: shuffle maximum, then one slot per warp
    for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    __shared__ float wmax[8];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = zmax;
    block_reduce_store<6, 256>(all, red, stats + (int64_t)blockIdx.x * 6);
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < 8; ++w) m = fmaxf(m, wmax[w]);
        stats[(int64_t)blockIdx.x * 6 + 5] = (double)m;
    }
}

cudaError_t launch_normals_hist(int rounds, uint64_t seed, int64_t n_quads, int n_steps, int n_bins, double lo, double hi,
                                unsigned long long* hist_dev, double* stats_dev, int grid, cudaStream_t s) {
    if (n_bins < 1 || n_bins > kNormalBinsMax || !(hi > lo)) return cudaErrorInvalidValue;
    const float inv_width = (float)((double)n_bins / (hi - lo));
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    if (rounds == 10) normals_hist_kernel<10><<<grid, 256, 0, s>>>(k0, k1, n_quads, n_steps, n_bins, (float)lo, inv_width, hist_dev, stats_dev);
    else if (rounds == 7) normals_hist_kernel<7><<<grid, 256, 0, s>>>(k0, k1, n_quads, n_steps, n_bins, (float)lo, inv_width, hist_dev, stats_dev);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K1z: block = 128 paths (one per thread); Z is consumed in tiles of kZSteps steps.  The tile of the NEXT step
// block is fetched with cp.async (LDGSTS, 16-byte chunks, coalesced along each path's row) into the second half
// of a double buffer while the current tile is turned into prices, so the row-major read of Z, the FP64 exp and
// the timestep-major write of S overlap instead of alternating.
constexpr int kZPaths = 128;
constexpr int kZSteps = 16;
constexpr int kZPitch = kZSteps + 2;        // doubles per shared row: 16-byte aligned rows, spreads the banks

__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename XT, bool WIDE>
__global__ void __launch_bounds__(kZPaths) normals_paths_kernel(const double* __restrict__ Z, XT* __restrict__ S,
                                                                int64_t ld, int n_steps, int64_t n_local,
                                                                GbmParams g) {
    __shared__ __align__(16) double tile[2][kZPaths][kZPitch];
    const int n_tiles = (n_steps + kZSteps - 1) / kZSteps;

    auto fetch = [&](int64_t pb, int k, int buf) {
        const int t0 = k * kZSteps;
        const int cols = min(kZSteps, n_steps - t0);
        if (WIDE) {
            // 16-byte chunks: chunk c of row r; consecutive threads walk along a row first (coalesced)
            constexpr int CH = kZSteps / 2;
            for (int idx = threadIdx.x; idx < kZPaths * CH; idx += kZPaths) {
                const int r = idx / CH, c = idx % CH;
                if (pb + r < n_local && 2 * c < cols)            // n_steps is even in WIDE mode: chunks are whole
                    cp_async_16(&tile[buf][r][2 * c], Z + (pb + r) * n_steps + t0 + 2 * c);
            }
        } else {
            for (int idx = threadIdx.x; idx < kZPaths * kZSteps; idx += kZPaths) {
                const int r = idx / kZSteps, c = idx % kZSteps;
                if (pb + r < n_local && c < cols) cp_async_8(&tile[buf][r][c], Z + (pb + r) * n_steps + t0 + c);
            }
        }
        cp_async_commit();
    };

    for (int64_t pb = (int64_t)blockIdx.x * kZPaths; pb < n_local; pb += (int64_t)gridDim.x * kZPaths) {
        const int64_t p = pb + threadIdx.x;
        double L = 0.0;
        if (p < n_local) S[p] = (XT)g.S0;
        fetch(pb, 0, 0);
        for (int k = 0; k < n_tiles; ++k) {
            const int buf = k & 1;
            if (k + 1 < n_tiles) {
                fetch(pb, k + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            if (p < n_local) {
                const int t0 = k * kZSteps;
                const int jmax = min(kZSteps, n_steps - t0);
#pragma unroll 4
                for (int j = 0; j < jmax; ++j) {
                    L += fma(g.vol, tile[buf][threadIdx.x][j], g.drift);
                    S[(int64_t)(t0 + j + 1) * ld + p] = (XT)(g.S0 * exp(L));
                }
            }
            __syncthreads();             // the buffer is refilled two iterations later
        }
    }
}

cudaError_t launch_from_normals(int dtype, const double* Z_dev, void* S, int64_t ld, int n_steps, int64_t n_local,
                                GbmParams g, cudaStream_t s) {
    int64_t blocks = (n_local + kZPaths - 1) / kZPaths;
    if (blocks > device_sm_count() * 32) blocks = device_sm_count() * 32;
    if (blocks < 1) blocks = 1;
    const bool wide = (n_steps % 2 == 0) && (((uintptr_t)Z_dev & 15) == 0);
    const int b = (int)blocks;
    if (dtype == 1) {
        if (wide) normals_paths_kernel<float, true><<<b, kZPaths, 0, s>>>(Z_dev, (float*)S, ld, n_steps, n_local, g);
        else normals_paths_kernel<float, false><<<b, kZPaths, 0, s>>>(Z_dev, (float*)S, ld, n_steps, n_local, g);
    } else {
        if (wide) normals_paths_kernel<double, true><<<b, kZPaths, 0, s>>>(Z_dev, (double*)S, ld, n_steps, n_local, g);
        else normals_paths_kernel<double, false><<<b, kZPaths, 0, s>>>(Z_dev, (double*)S, ld, n_steps, n_local, g);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Reference layout S[p][t] (row-major, f64) -> timestep-major storage.  32 x 32 tiles.
template <typename XT>
__global__ void __launch_bounds__(256) transpose_in_kernel(const double* __restrict__ in, XT* __restrict__ out,
                                                           int64_t ld, int n_cols, int64_t n_local) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
    const int64_t n_ptiles = (n_local + 31) / 32;
    const int n_ttiles = (n_cols + 31) / 32;
    for (int64_t tileid = blockIdx.x; tileid < n_ptiles * n_ttiles; tileid += gridDim.x) {
        const int64_t pt = tileid / n_ttiles;
        const int tt = (int)(tileid % n_ttiles);
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int64_t p = pt * 32 + r;
            const int t = tt * 32 + tx;
            if (p < n_local && t < n_cols) tile[r][tx] = __ldg(in + p * n_cols + t);
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int t = tt * 32 + r;
            const int64_t p = pt * 32 + tx;
            if (p < n_local && t < n_cols) out[(int64_t)t * ld + p] = (XT)tile[tx][r];
        }
    }
}

cudaError_t launch_transpose_in(int dtype, const double* in, void* S, int64_t ld, int n_cols, int64_t n_local,
                                cudaStream_t s) {
    int64_t tiles = ((n_local + 31) / 32) * ((n_cols + 31) / 32);
    if (tiles > device_sm_count() * 32) tiles = device_sm_count() * 32;
    if (tiles < 1) tiles = 1;
    if (dtype == 1)
        transpose_in_kernel<float><<<(int)tiles, 256, 0, s>>>(in, (float*)S, ld, n_cols, n_local);
    else
        transpose_in_kernel<double><<<(int)tiles, 256, 0, s>>>(in, (double*)S, ld, n_cols, n_local);
    return cudaGetLastError();
}

// rows [p0, p1) back to the reference layout out[p - p0][t] (f64)
template <typename XT>
__global__ void gather_rows_kernel(const XT* __restrict__ S, int64_t ld, int n_cols, int64_t p0, int64_t p1,
                                   double* __restrict__ out) {
    const int64_t total = (p1 - p0) * n_cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + i / n_cols;
        const int t = (int)(i % n_cols);
        out[i] = (double)S[(int64_t)t * ld + p];
    }
}

cudaError_t launch_gather_rows(int dtype, const void* S, int64_t ld, int n_cols, int64_t p0, int64_t p1,
                               double* out_dev, cudaStream_t s) {
    int64_t total = (p1 - p0) * n_cols;
    int64_t blocks = (total + 255) / 256;
    if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
    if (blocks < 1) blocks = 1;
    if (dtype == 1)
        gather_rows_kernel<float><<<(int)blocks, 256, 0, s>>>((const float*)S, ld, n_cols, p0, p1, out_dev);
    else
        gather_rows_kernel<double><<<(int)blocks, 256, 0, s>>>((const double*)S, ld, n_cols, p0, p1, out_dev);
    return cudaGetLastError();
}

// out[p] = S[steps[p]][p]: each path's value at its own step (e.g. its exercise step)
template <typename XT>
__global__ void gather_steps_kernel(const XT* __restrict__ S, int64_t ld, int n_cols, int64_t n,
                                    const int32_t* __restrict__ steps, double* __restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        int t = steps[p];
        t = t < 0 ? 0 : (t >= n_cols ? n_cols - 1 : t);
        out[p] = (double)S[(int64_t)t * ld + p];
    }
}

cudaError_t launch_gather_steps(int dtype, const void* S, int64_t ld, int n_cols, int64_t n, const int32_t* steps_dev,
                                double* out_dev, cudaStream_t s) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
    if (blocks < 1) blocks = 1;
    if (dtype == 1)
        gather_steps_kernel<float><<<(int)blocks, 256, 0, s>>>((const float*)S, ld, n_cols, n, steps_dev, out_dev);
    else
        gather_steps_kernel<double><<<(int)blocks, 256, 0, s>>>((const double*)S, ld, n_cols, n, steps_dev, out_dev);
    return cudaGetLastError();
}

template <typename XT>
__global__ void column_to_f64_kernel(const XT* __restrict__ col, int64_t n, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)col[i];
}

cudaError_t launch_column_to_f64(int dtype, const void* col, int64_t n, double* out_dev, cudaStream_t s) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
    if (blocks < 1) blocks = 1;
    if (dtype == 1)
        column_to_f64_kernel<float><<<(int)blocks, 256, 0, s>>>((const float*)col, n, out_dev);
    else
        column_to_f64_kernel<double><<<(int)blocks, 256, 0, s>>>((const double*)col, n, out_dev);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Column statistics for adopted path matrices: grid = (n_chunks, n_cols); shifted sums (shift = first element
// of the column) so the variance does not cancel when the spread is small against the level.  Partials are
// finished on the host in a fixed order -> deterministic maps.
template <typename XT>
__global__ void __launch_bounds__(256) column_stats_kernel(const XT* __restrict__ S, int64_t ld, int64_t n_local,
                                                           int n_chunks, double* __restrict__ partial,
                                                           double* __restrict__ shift) {
    __shared__ double red[8 * 2];
    const int col = blockIdx.y, chunk = blockIdx.x;
    const XT* x = S + (int64_t)col * ld;
    const double c = (double)x[0];
    if (chunk == 0 && threadIdx.x == 0) shift[col] = c;
    const int64_t per = (n_local + n_chunks - 1) / n_chunks;
    const int64_t lo = (int64_t)chunk * per;
    const int64_t hi = (lo + per < n_local) ? lo + per : n_local;
    double acc[2] = {0.0, 0.0};
    for (int64_t p = lo + threadIdx.x; p < hi; p += blockDim.x) {
        const double dlt = (double)x[p] - c;
        acc[0] += dlt;
        acc[1] = fma(dlt, dlt, acc[1]);
    }
    block_reduce_store<2, 256>(acc, red, partial + ((int64_t)col * n_chunks + chunk) * 2);
}

cudaError_t launch_column_stats(int dtype, const void* S, int64_t ld, int n_cols, int64_t n_local, int n_chunks,
                                double* partial_dev, double* shift_dev, cudaStream_t s) {
    dim3 grid(n_chunks, n_cols);
    if (dtype == 1)
        column_stats_kernel<float><<<grid, 256, 0, s>>>((const float*)S, ld, n_local, n_chunks, partial_dev, shift_dev);
    else
        column_stats_kernel<double><<<grid, 256, 0, s>>>((const double*)S, ld, n_local, n_chunks, partial_dev, shift_dev);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Down-and-in barrier: precompute_barrier_hit_matrix (amc.py:171-176) is a running OR over time of
// S <= barrier, i.e. a [P, n+1] bool matrix; one int per path (first step at which the path is knocked in,
// n+1 = never) carries the same information.
template <typename XT>
__global__ void first_hit_kernel(const XT* __restrict__ S, int64_t ld, int n_cols, int64_t n_local, double barrier,
                                 int32_t* __restrict__ first_hit) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_local; p += (int64_t)gridDim.x * blockDim.x) {
        int fh = n_cols;
        for (int t = 0; t < n_cols; ++t) {
            const double x = (double)S[(int64_t)t * ld + p];
            if (x <= barrier && fh == n_cols) fh = t;
        }
        first_hit[p] = fh;
    }
}

cudaError_t launch_first_hit(int dtype, const void* S, int64_t ld, int n_cols, int64_t n_local, double barrier,
                             int32_t* first_hit_dev, cudaStream_t s) {
    int64_t blocks = (n_local + 255) / 256;
    if (blocks > device_sm_count() * 8) blocks = device_sm_count() * 8;
    if (blocks < 1) blocks = 1;
    if (dtype == 1)
        first_hit_kernel<float><<<(int)blocks, 256, 0, s>>>((const float*)S, ld, n_cols, n_local, barrier, first_hit_dev);
    else
        first_hit_kernel<double><<<(int)blocks, 256, 0, s>>>((const double*)S, ld, n_cols, n_local, barrier, first_hit_dev);
    return cudaGetLastError();
}

__global__ void hit_matrix_kernel(const int32_t* __restrict__ first_hit, int n_cols, int64_t n_local,
                                  uint8_t* __restrict__ out) {
    const int64_t total = n_local * n_cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / n_cols;
        const int t = (int)(i % n_cols);
        out[i] = (t >= first_hit[p]) ? 1 : 0;
    }
}

cudaError_t launch_hit_matrix(const int32_t* first_hit_dev, int n_cols, int64_t n_local, uint8_t* out_dev,
                              cudaStream_t s) {
    int64_t total = n_local * n_cols;
    int64_t blocks = (total + 255) / 256;
    if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
    if (blocks < 1) blocks = 1;
    hit_matrix_kernel<<<(int)blocks, 256, 0, s>>>(first_hit_dev, n_cols, n_local, out_dev);
    return cudaGetLastError();
}

}  // namespace amc
