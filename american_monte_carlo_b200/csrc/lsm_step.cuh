// LSM backward-induction kernels for sm_100a.
//
//  lsm_step_kernel<XT, D>  one launch per time step t: fused
//        (1) exercise decision at step t for every path  (amc.py:147-149 at maturity, :154-162/:90-94 below it)
//        (2) moment sums of the regression of step t-1   (the O(P) part of amc.py:110-128)
//      Reads column t, column t-1 and the per-path state once, writes the state once:
//      2*b_S + 16 algorithmic bytes per path-step -- an HBM-streaming kernel (FP64 FMA work ~30-90 flop per
//      32-48 B, far below the tensor-core regime; no dense contraction exists at k <= 11).
//  lsm_solve_kernel        single block between two step launches: fixed-order reduction of the per-block
//      partial sums, then one thread runs lsm_solve.h (Cholesky + change of basis + Jacobi SVD + numpy's
//      rank rule) and leaves the continuation polynomial in device memory for the next step launch.
//
// State: U[p] = cashflow of path p discounted to time 0 (= cashflows * exp(-r dt exercise_times) of
// amc.py:128,196, which the reference recomputes at every step).  The regression target at step t is
// Y = U * exp(r dt t); the scalar factor is applied to the reduced sums, not per path.
//
// This header holds the step kernels and their templated launchers; it is compiled once per path storage type
// (lsm_step_f32.cu, lsm_step_f64.cu) so the two sets of 11 degree instantiations build in parallel.
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"

namespace amc {

template <typename XT> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// power sums of one path: acc[m-1] += z^m (m = 1..2D), acc[2D+m] += z^m * y (m = 0..D)
template <int D>
__device__ __forceinline__ void accumulate_moments(double z, double y, double (&acc)[3 * D + 1]) {
    acc[2 * D] += y;
    double p = 1.0;
#pragma unroll
    for (int m = 1; m <= 2 * D; ++m) {
        p *= z;
        acc[m - 1] += p;
        if (m <= D) acc[2 * D + m] = fma(p, y, acc[2 * D + m]);
    }
}

template <int D>
__device__ __forceinline__ double horner(const double (&gam)[D + 1], double z) {
    double f = gam[D];
#pragma unroll
    for (int m = D - 1; m >= 0; --m) f = fma(f, z, gam[m]);
    return f;
}

// One path: decision at t_dec, then moments at t_dec-1.  All flags are launch-uniform.
template <int D>
__device__ __forceinline__ bool path_step(const StepArgs& a, const double (&gam)[D + 1], double xd, double xr,
                                          double& u, int& tau, int fh, double (&acc)[3 * D + 1]) {
    bool changed = false;
    if (a.mode != kObserve) {
        const double iv = a.is_put ? (a.K - xd) : (xd - a.K);
        const bool in = (fh <= a.t_dec);
        if (a.mode == kMaturity) {
            // cashflows[hit] = max(payoff, 0), exercise_times[hit] = n; everything else stays 0 / n
            u = (in && iv > 0.0) ? iv * a.disc_dec : 0.0;
            tau = a.t_dec;
            changed = true;
        } else {
            const double zd = fma(xd, a.isg_dec, -a.mu_dec * a.isg_dec);
            const double fit = horner<D>(gam, zd);
            // candidates: knocked in AND in the money; exercise iff payoff > max(fit, 0)  (strict)
            if (in && iv > 0.0 && iv > fit) {
                u = iv * a.disc_dec;
                tau = a.t_dec;
                changed = true;
            }
        }
    }
    if (a.moments) {
        const double zr = fma(xr, a.isg_reg, -a.mu_reg * a.isg_reg);
        accumulate_moments<D>(zr, u, acc);
    } else {
        acc[2 * D] += u;
    }
    return changed;
}

// ---------------------------------------------------------------------------------------------------------
// Hot-path arithmetic with every launch-uniform flag folded into constants (American decision + moments, no
// barrier, no exercise-step array, full tile): ~24 FP64 instructions per path at degree 3.
struct FastConsts {
    double sgn, sgnK;       // payoff = fma(sgn, x, sgnK): put -> K - x, call -> x - K
    double da, db;          // z_dec = fma(x, da, db)   (= (x - mu) * isg up to one rounding)
    double ra, rb;          // z_reg = fma(x, ra, rb)
    double disc;
};

// acc[m-1] += z^m (m = 1..2D), acc[2D+m] += z^m y (m = 0..D) with D-1 multiplies: the high powers are formed
// inside the accumulating FMA as z^(m-D) * z^D.
template <int D>
__device__ __forceinline__ void accumulate_moments_fast(double z, double y, bool with_h, double (&acc)[3 * D + 1]) {
    acc[2 * D] += y;
    if (D == 0) return;
    double p[D + 1];
    p[0] = 1.0;
    p[1] = z;
#pragma unroll
    for (int m = 2; m <= D; ++m) p[m] = p[m - 1] * z;
#pragma unroll
    for (int m = 1; m <= D; ++m) acc[2 * D + m] = fma(p[m], y, acc[2 * D + m]);
    if (with_h) {              // block-uniform: in a contract batch only contract 0 sums the powers of the column
#pragma unroll
        for (int m = 1; m <= D; ++m) {
            acc[m - 1] += p[m];
            acc[D + m - 1] = fma(p[m], p[D], acc[D + m - 1]);
        }
    }
}

template <int D>
__device__ __forceinline__ bool fast_path_step(const FastConsts& c, const double (&gam)[D + 1], double xd, double xr,
                                               double& u, bool with_h, double (&acc)[3 * D + 1]) {
    const double iv = fma(c.sgn, xd, c.sgnK);
    const double zd = fma(xd, c.da, c.db);
    const double fit = horner<D>(gam, zd);
    const bool ex = (iv > 0.0) && (iv > fit);
    if (ex) u = iv * c.disc;
    accumulate_moments_fast<D>(fma(xr, c.ra, c.rb), u, with_h, acc);
    return ex;
}

// ---------------------------------------------------------------------------------------------------------
// TMA-pipelined variant of the step kernel (the default).  Each persistent block owns a ring of kStages shared-
// memory stages; one elected thread issues 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) of the next
// tiles of S_t, S_{t-1} and U while all 8 warps compute on the current tile, completion signalled through
// mbarriers (complete_tx).  Loads therefore live in shared memory instead of registers: the bytes in flight per
// SM are set by the ring (kStages x 24 KB at f64), not by occupancy x registers, which is what limited the
// register-staged kernel above to 0.79 of the copy roofline (ncu: 108 registers, 25 % occupancy).
// The updated state goes straight from registers to global memory (coalesced 16-byte stores).
constexpr int kTile = 1024;       // paths per tile
constexpr int kStages = 4;

// updated state of two adjacent paths: one 16-byte (f64 state) or 8-byte (f32 state) store
__device__ __forceinline__ void store_state2(double* dst, double2 u, uint64_t policy) {
    st_hint(reinterpret_cast<double2*>(dst), u, policy);
}
__device__ __forceinline__ void store_state2(float* dst, double2 u, uint64_t policy) {
    st_hint(reinterpret_cast<float2*>(dst), make_float2((float)u.x, (float)u.y), policy);
}

template <typename XT, typename UT>
struct StageBytes { static constexpr int value = kTile * (2 * (int)sizeof(XT) + (int)sizeof(UT)); };

template <typename XT, typename UT, int D>
__global__ void __launch_bounds__(kStepThreads) lsm_step_tma_kernel(const StepArgs a_in) {
    constexpr int NACC = 3 * D + 1;
    // contract batches: this block's contract = blockIdx.y (block-uniform overrides of the launch arguments)
    StepArgs a = a_in;
    if (a_in.batch) {
        const BatchContract bc = a_in.batch[blockIdx.y];
        a.K = bc.K;
        a.is_put = bc.is_put;
        if (a.mode != kMaturity) a.mode = bc.is_american ? kDecide : kObserve;
        a.U = static_cast<UT*>(a_in.U) + (int64_t)blockIdx.y * a_in.u_stride;
        a.coef = a_in.coef + (int64_t)blockIdx.y * a_in.coef_stride;
        a.partials = a_in.partials + (int64_t)blockIdx.y * gridDim.x * kAccStride;
    }
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ double red[(kStepThreads / 32) * NACC];
    __shared__ uint64_t full[kStages];

    UT* const Ug = static_cast<UT*>(a.U);
    using U2 = typename Vec2<UT>::type;

    pdl_launch_dependents();     // let the solve kernel of this step become resident right away

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
    const XT* xdec = static_cast<const XT*>(a.x_dec);
    const XT* xreg = static_cast<const XT*>(a.x_reg);
    const bool need_dec = (a.mode != kObserve);
    const bool need_u_in = (a.mode != kMaturity);
    const bool write_u = (a.mode != kObserve);

    const int64_t n_tiles = (a.n_paths + kTile - 1) / kTile;
    const int my_tiles = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles blockIdx.x + i*grid
    // this block's tiles as a running path offset: first + i * stride (stride < 0 when the launch walks the column
    // from the high end, see StepArgs::reverse)
    const int64_t p_first = (a.reverse ? (n_tiles - 1 - (int64_t)blockIdx.x) : (int64_t)blockIdx.x) * kTile;
    const int64_t p_stride = (a.reverse ? -(int64_t)gridDim.x : (int64_t)gridDim.x) * kTile;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // L2 is managed explicitly: S_t is dead after this launch (evict_first); S_{t-1} and the state are re-read by the
    // next launch (evict_last).  A/B of the policies: profiles/r1c_l2_ab.txt.
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    // thread 0 only: start the copies of the tile at path offset p0 into stage s.  parts bit 0: arm the barrier with
    // the tile's total byte count and copy the two path columns; bit 1: copy the state.  (Split because the columns
    // are immutable and may be fetched before the programmatic-dependency wait, the state may not.)
    auto issue = [&](int64_t p0, int s, int parts) {
        int64_t valid = a.n_paths - p0;
        if (valid > kTile) valid = kTile;
        const uint32_t elems = (uint32_t)((valid + 31) / 32 * 32);     // columns are padded to 32 elements
        unsigned char* st = ring + (size_t)s * StageBytes<XT, UT>::value;
        const uint32_t bx = elems * (uint32_t)sizeof(XT), bu = elems * (uint32_t)sizeof(UT);
        if (parts & 1) {
            mbar_expect_tx(&full[s], (need_dec ? bx : 0u) + (a.moments ? bx : 0u) + (need_u_in ? bu : 0u));
            if (need_dec) tma_load_1d_hint(st, xdec + p0, bx, &full[s], pol_stream);
            if (a.moments) tma_load_1d_hint(st + kTile * sizeof(XT), xreg + p0, bx, &full[s], pol_keep);
        }
        if ((parts & 2) && need_u_in) tma_load_1d_hint(st + 2 * kTile * sizeof(XT), Ug + p0, bu, &full[s], pol_keep);
    };

    // prologue: the path columns of the first tiles are on their way while the previous kernel of the chain (the
    // solve of this step) is still finishing; the state and the polynomial are touched only after the wait
    int64_t p_issue = p_first;                           // thread 0: offset of the next tile to be issued
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages - 1 && i < my_tiles; ++i) issue(p_first + i * p_stride, i, 1);
    }
    pdl_wait();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages - 1 && i < my_tiles; ++i) { issue(p_issue, i, 2); p_issue += p_stride; }
    }
    double gam[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) gam[i] = (a.mode == kDecide) ? a.coef[i] : 0.0;

    using V2 = typename Vec2<XT>::type;
    const bool fast_ok = (a.mode == kDecide) && a.moments && !a.first_hit && !a.tau;
    FastConsts fc;
    fc.sgn = a.is_put ? -1.0 : 1.0;
    fc.sgnK = a.is_put ? a.K : -a.K;
    fc.da = a.isg_dec; fc.db = -a.mu_dec * a.isg_dec;
    fc.ra = a.isg_reg; fc.rb = -a.mu_reg * a.isg_reg;
    fc.disc = a.disc_dec;
    // the power sums of the regressed column are the same for every contract of a batch: contract 0 computes them
    // (the solve kernel reads them from its rows), the fast path of the others only forms the cross sums with y
    const bool with_h = !a_in.batch || blockIdx.y == 0;
    constexpr int NK = kTile / 2 / kStepThreads;
    int64_t p0 = p_first;
    int s = 0, s_issue = (kStages - 1) % kStages;
    uint32_t phase = 0;
    for (int i = 0; i < my_tiles; ++i, p0 += p_stride) {
        if (threadIdx.x == 0 && i + kStages - 1 < my_tiles) {
            issue(p_issue, s_issue, 3);
            p_issue += p_stride;
        }
        s_issue = (s_issue + 1 == kStages) ? 0 : s_issue + 1;
        mbar_wait(&full[s], phase);

        const unsigned char* st = ring + (size_t)s * StageBytes<XT, UT>::value;
        const V2* sxd = reinterpret_cast<const V2*>(st) + threadIdx.x;
        const V2* sxr = reinterpret_cast<const V2*>(st + kTile * sizeof(XT)) + threadIdx.x;
        const U2* su = reinterpret_cast<const U2*>(st + 2 * kTile * sizeof(XT)) + threadIdx.x;

        if (fast_ok && p0 + kTile <= a.n_paths) {
            UT* const up = Ug + p0 + 2 * threadIdx.x;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const V2 vd = sxd[k * kStepThreads], vr = sxr[k * kStepThreads];
                const U2 uv = su[k * kStepThreads];
                double2 u = make_double2((double)uv.x, (double)uv.y);
                bool changed = fast_path_step<D>(fc, gam, (double)vd.x, (double)vr.x, u.x, with_h, acc);
                changed |= fast_path_step<D>(fc, gam, (double)vd.y, (double)vr.y, u.y, with_h, acc);
                // the state is written only where a path exercised (one vector store per pair): below maturity most
                // pairs are untouched, which removes most of the write traffic
                if (changed) store_state2(up + 2 * k * kStepThreads, u, pol_keep);
            }
        } else {
            int64_t valid64 = a.n_paths - p0;
            const int valid = (int)(valid64 > kTile ? kTile : valid64);
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const int j = threadIdx.x + k * kStepThreads;        // pair index inside the tile
                const int e0 = 2 * j;
                if (e0 < valid) {
                    const bool two = (e0 + 1 < valid);
                    double xd0 = 0, xd1 = 0, xr0 = 0, xr1 = 0;
                    double2 u = make_double2(0.0, 0.0);
                    int2 f = make_int2(0, 0), t = make_int2(0, 0);
                    if (need_dec) { const V2 v = sxd[k * kStepThreads]; xd0 = (double)v.x; xd1 = (double)v.y; }
                    if (a.moments) { const V2 v = sxr[k * kStepThreads]; xr0 = (double)v.x; xr1 = (double)v.y; }
                    if (need_u_in) { const U2 uv = su[k * kStepThreads]; u = make_double2((double)uv.x, (double)uv.y); }
                    const int64_t p = p0 + e0;
                    if (a.first_hit) { f.x = __ldg(a.first_hit + p); if (two) f.y = __ldg(a.first_hit + p + 1); }
                    if (a.tau && need_u_in) { t.x = a.tau[p]; if (two) t.y = a.tau[p + 1]; }
                    bool changed = path_step<D>(a, gam, xd0, xr0, u.x, t.x, f.x, acc);
                    if (two) changed |= path_step<D>(a, gam, xd1, xr1, u.y, t.y, f.y, acc);
                    if (write_u && changed) {
                        if (two) {
                            store_state2(Ug + p, u, pol_keep);
                            if (a.tau) *reinterpret_cast<int2*>(a.tau + p) = t;
                        } else {
                            Ug[p] = (UT)u.x;
                            if (a.tau) a.tau[p] = t.x;
                        }
                    }
                }
            }
        }
        __syncthreads();                 // every warp is done with stage s before it is refilled
        if (++s == kStages) { s = 0; phase ^= 1u; }
    }
    block_reduce_store<NACC, kStepThreads, kAccStride>(acc, red, a.partials + (int64_t)blockIdx.x * kAccStride);
}

// ---------------------------------------------------------------------------------------------------------
// launchers
template <typename XT, typename UT, int D>
static cudaError_t launch_step_t(int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch) {
    constexpr int smem = kStages * StageBytes<XT, UT>::value;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lsm_step_tma_kernel<XT, UT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    return launch_ex(lsm_step_tma_kernel<XT, UT, D>, dim3(grid, n_batch), kStepThreads, smem, s, pdl, a);
}

template <typename XT, typename UT>
static cudaError_t launch_step_d(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch) {
    switch (degree) {
        case 0: return launch_step_t<XT, UT, 0>(grid, a, s, pdl, n_batch);
        case 1: return launch_step_t<XT, UT, 1>(grid, a, s, pdl, n_batch);
        case 2: return launch_step_t<XT, UT, 2>(grid, a, s, pdl, n_batch);
        case 3: return launch_step_t<XT, UT, 3>(grid, a, s, pdl, n_batch);
        case 4: return launch_step_t<XT, UT, 4>(grid, a, s, pdl, n_batch);
        case 5: return launch_step_t<XT, UT, 5>(grid, a, s, pdl, n_batch);
        case 6: return launch_step_t<XT, UT, 6>(grid, a, s, pdl, n_batch);
        case 7: return launch_step_t<XT, UT, 7>(grid, a, s, pdl, n_batch);
        case 8: return launch_step_t<XT, UT, 8>(grid, a, s, pdl, n_batch);
        case 9: return launch_step_t<XT, UT, 9>(grid, a, s, pdl, n_batch);
        case 10: return launch_step_t<XT, UT, 10>(grid, a, s, pdl, n_batch);
    }
    return cudaErrorInvalidValue;
}

template <typename XT, typename UT, int D>
static int occupancy_blocks() {
    int nb = 0;
    constexpr int smem = kStages * StageBytes<XT, UT>::value;
    cudaFuncSetAttribute(lsm_step_tma_kernel<XT, UT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lsm_step_tma_kernel<XT, UT, D>, kStepThreads, smem);
    return nb;
}

template <typename XT, typename UT>
static int occupancy_d(int degree) {
    switch (degree) {
        case 0: return occupancy_blocks<XT, UT, 0>();
        case 1: return occupancy_blocks<XT, UT, 1>();
        case 2: return occupancy_blocks<XT, UT, 2>();
        case 3: return occupancy_blocks<XT, UT, 3>();
        case 4: return occupancy_blocks<XT, UT, 4>();
        case 5: return occupancy_blocks<XT, UT, 5>();
        case 6: return occupancy_blocks<XT, UT, 6>();
        case 7: return occupancy_blocks<XT, UT, 7>();
        case 8: return occupancy_blocks<XT, UT, 8>();
        case 9: return occupancy_blocks<XT, UT, 9>();
        case 10: return occupancy_blocks<XT, UT, 10>();
    }
    return 1;
}

}  // namespace amc
