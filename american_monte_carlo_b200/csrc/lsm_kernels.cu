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
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace amc {

template <typename XT> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename XT>
__device__ __forceinline__ void load_pair(const XT* col, int64_t unit, double& a, double& b) {
    typename Vec2<XT>::type v = __ldg(reinterpret_cast<const typename Vec2<XT>::type*>(col) + unit);
    a = (double)v.x;
    b = (double)v.y;
}

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

template <typename XT, int D>
__global__ void __launch_bounds__(kStepThreads) lsm_step_kernel(const StepArgs a) {
    constexpr int NACC = 3 * D + 1;
    __shared__ double red[(kStepThreads / 32) * NACC];
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
    double gam[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) gam[i] = (a.mode == kDecide) ? a.coef[i] : 0.0;

    const XT* xdec = static_cast<const XT*>(a.x_dec);
    const XT* xreg = static_cast<const XT*>(a.x_reg);
    const bool need_dec = (a.mode != kObserve);
    const bool need_u_in = (a.mode != kMaturity);
    const bool write_u = (a.mode != kObserve);

    const int64_t n_units = a.n_paths >> 1;                       // full pairs
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;

    double2* U2 = reinterpret_cast<double2*>(a.U);
    int2* T2 = reinterpret_cast<int2*>(a.tau);
    const int2* F2 = reinterpret_cast<const int2*>(a.first_hit);

    // two independent pairs per iteration: all loads are issued before the arithmetic
    int64_t q = tid;
    for (; q + stride < n_units; q += 2 * stride) {
        const int64_t q1 = q + stride;
        double xd0 = 0, xd1 = 0, xd2 = 0, xd3 = 0, xr0 = 0, xr1 = 0, xr2 = 0, xr3 = 0;
        double2 u0 = make_double2(0.0, 0.0), u1 = make_double2(0.0, 0.0);
        int2 f0 = make_int2(0, 0), f1 = make_int2(0, 0), t0 = make_int2(0, 0), t1 = make_int2(0, 0);
        if (need_dec) { load_pair<XT>(xdec, q, xd0, xd1); load_pair<XT>(xdec, q1, xd2, xd3); }
        if (a.moments) { load_pair<XT>(xreg, q, xr0, xr1); load_pair<XT>(xreg, q1, xr2, xr3); }
        if (need_u_in) { u0 = U2[q]; u1 = U2[q1]; }
        if (F2) { f0 = __ldg(F2 + q); f1 = __ldg(F2 + q1); }
        if (T2 && need_u_in) { t0 = T2[q]; t1 = T2[q1]; }
        path_step<D>(a, gam, xd0, xr0, u0.x, t0.x, f0.x, acc);
        path_step<D>(a, gam, xd1, xr1, u0.y, t0.y, f0.y, acc);
        path_step<D>(a, gam, xd2, xr2, u1.x, t1.x, f1.x, acc);
        path_step<D>(a, gam, xd3, xr3, u1.y, t1.y, f1.y, acc);
        if (write_u) {
            U2[q] = u0; U2[q1] = u1;
            if (T2) { T2[q] = t0; T2[q1] = t1; }
        }
    }
    for (; q < n_units; q += stride) {
        double xd0 = 0, xd1 = 0, xr0 = 0, xr1 = 0;
        double2 u0 = make_double2(0.0, 0.0);
        int2 f0 = make_int2(0, 0), t0 = make_int2(0, 0);
        if (need_dec) load_pair<XT>(xdec, q, xd0, xd1);
        if (a.moments) load_pair<XT>(xreg, q, xr0, xr1);
        if (need_u_in) u0 = U2[q];
        if (F2) f0 = __ldg(F2 + q);
        if (T2 && need_u_in) t0 = T2[q];
        path_step<D>(a, gam, xd0, xr0, u0.x, t0.x, f0.x, acc);
        path_step<D>(a, gam, xd1, xr1, u0.y, t0.y, f0.y, acc);
        if (write_u) {
            U2[q] = u0;
            if (T2) T2[q] = t0;
        }
    }
    // odd path count: the last path is handled by the thread that would own the next unit
    if ((a.n_paths & 1) && tid == (n_units % stride)) {
        const int64_t p = a.n_paths - 1;
        double xd = need_dec ? (double)xdec[p] : 0.0;
        double xr = a.moments ? (double)xreg[p] : 0.0;
        double u = need_u_in ? a.U[p] : 0.0;
        int fh = a.first_hit ? a.first_hit[p] : 0;
        int tau = (a.tau && need_u_in) ? a.tau[p] : 0;
        path_step<D>(a, gam, xd, xr, u, tau, fh, acc);
        if (write_u) {
            a.U[p] = u;
            if (a.tau) a.tau[p] = tau;
        }
    }
    block_reduce_store<NACC, kStepThreads, kAccStride>(acc, red, a.partials + (int64_t)blockIdx.x * kAccStride);
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
__device__ __forceinline__ void accumulate_moments_fast(double z, double y, double (&acc)[3 * D + 1]) {
    acc[2 * D] += y;
    if (D == 0) return;
    double p[D + 1];
    p[0] = 1.0;
    p[1] = z;
#pragma unroll
    for (int m = 2; m <= D; ++m) p[m] = p[m - 1] * z;
#pragma unroll
    for (int m = 1; m <= D; ++m) {
        acc[m - 1] += p[m];
        acc[2 * D + m] = fma(p[m], y, acc[2 * D + m]);
        acc[D + m - 1] = fma(p[m], p[D], acc[D + m - 1]);
    }
}

template <int D>
__device__ __forceinline__ bool fast_path_step(const FastConsts& c, const double (&gam)[D + 1], double xd, double xr,
                                               double& u, double (&acc)[3 * D + 1]) {
    const double iv = fma(c.sgn, xd, c.sgnK);
    const double zd = fma(xd, c.da, c.db);
    const double fit = horner<D>(gam, zd);
    const bool ex = (iv > 0.0) && (iv > fit);
    if (ex) u = iv * c.disc;
    accumulate_moments_fast<D>(fma(xr, c.ra, c.rb), u, acc);
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

template <typename XT>
struct StageBytes { static constexpr int value = kTile * (2 * (int)sizeof(XT) + 8); };

template <typename XT, int D>
__global__ void __launch_bounds__(kStepThreads) lsm_step_tma_kernel(const StepArgs a) {
    constexpr int NACC = 3 * D + 1;
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ double red[(kStepThreads / 32) * NACC];
    __shared__ uint64_t full[kStages];

    pdl_launch_dependents();     // let the solve kernel of this step become resident right away
    pdl_wait();                  // ... and do not touch memory before the previous solve has finished

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
    double gam[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) gam[i] = (a.mode == kDecide) ? a.coef[i] : 0.0;

    const XT* xdec = static_cast<const XT*>(a.x_dec);
    const XT* xreg = static_cast<const XT*>(a.x_reg);
    const bool need_dec = (a.mode != kObserve);
    const bool need_u_in = (a.mode != kMaturity);
    const bool write_u = (a.mode != kObserve);

    const int64_t n_tiles = (a.n_paths + kTile - 1) / kTile;
    const int my_tiles = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles blockIdx.x + i*grid

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    auto tile_of = [&](int i) -> int64_t {
        const int64_t fwd = blockIdx.x + (int64_t)i * gridDim.x;
        return a.reverse ? (n_tiles - 1 - fwd) : fwd;
    };
    auto issue = [&](int i) {           // thread 0 only: start the copies of this block's i-th tile
        const int64_t tile = tile_of(i);
        const int64_t p0 = tile * kTile;
        int64_t valid = a.n_paths - p0;
        if (valid > kTile) valid = kTile;
        const uint32_t elems = (uint32_t)((valid + 31) / 32 * 32);     // columns are padded to 32 elements
        const int s = i % kStages;
        unsigned char* st = ring + (size_t)s * StageBytes<XT>::value;
        const uint32_t bx = elems * (uint32_t)sizeof(XT), bu = elems * 8u;
        const uint32_t total = (need_dec ? bx : 0u) + (a.moments ? bx : 0u) + (need_u_in ? bu : 0u);
        mbar_expect_tx(&full[s], total);
        if (a.l2_hints) {
            // S_t is dead after this launch; S_{t-1} and U are re-read by the next launch
            if (need_dec) tma_load_1d_hint(st, xdec + p0, bx, &full[s], pol_stream);
            if (a.moments) tma_load_1d_hint(st + kTile * sizeof(XT), xreg + p0, bx, &full[s], pol_keep);
            if (need_u_in) tma_load_1d_hint(st + 2 * kTile * sizeof(XT), a.U + p0, bu, &full[s], pol_keep);
        } else {
            if (need_dec) tma_load_1d(st, xdec + p0, bx, &full[s]);
            if (a.moments) tma_load_1d(st + kTile * sizeof(XT), xreg + p0, bx, &full[s]);
            if (need_u_in) tma_load_1d(st + 2 * kTile * sizeof(XT), a.U + p0, bu, &full[s]);
        }
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages - 1 && i < my_tiles; ++i) issue(i);
    }

    using V2 = typename Vec2<XT>::type;
    const bool fast_ok = (a.mode == kDecide) && a.moments && !a.first_hit && !a.tau;
    FastConsts fc;
    fc.sgn = a.is_put ? -1.0 : 1.0;
    fc.sgnK = a.is_put ? a.K : -a.K;
    fc.da = a.isg_dec; fc.db = -a.mu_dec * a.isg_dec;
    fc.ra = a.isg_reg; fc.rb = -a.mu_reg * a.isg_reg;
    fc.disc = a.disc_dec;
    for (int i = 0; i < my_tiles; ++i) {
        const int s = i % kStages;
        if (threadIdx.x == 0 && i + kStages - 1 < my_tiles) issue(i + kStages - 1);
        mbar_wait(&full[s], (uint32_t)((i / kStages) & 1));

        const int64_t tile = tile_of(i);
        const int64_t p0 = tile * kTile;
        int64_t valid64 = a.n_paths - p0;
        const int valid = (int)(valid64 > kTile ? kTile : valid64);
        const unsigned char* st = ring + (size_t)s * StageBytes<XT>::value;
        const V2* sxd = reinterpret_cast<const V2*>(st);
        const V2* sxr = reinterpret_cast<const V2*>(st + kTile * sizeof(XT));
        const double2* su = reinterpret_cast<const double2*>(st + 2 * kTile * sizeof(XT));

        if (fast_ok && valid == kTile) {
#pragma unroll
            for (int k = 0; k < kTile / 2 / kStepThreads; ++k) {
                const int j = threadIdx.x + k * kStepThreads;
                const V2 vd = sxd[j], vr = sxr[j];
                double2 u = su[j];
                bool changed = fast_path_step<D>(fc, gam, (double)vd.x, (double)vr.x, u.x, acc);
                changed |= fast_path_step<D>(fc, gam, (double)vd.y, (double)vr.y, u.y, acc);
                if (changed) {
                    double2* dst = reinterpret_cast<double2*>(a.U + p0 + 2 * j);
                    if (a.l2_hints) st_hint(dst, u, pol_keep);
                    else *dst = u;
                }
            }
        } else
#pragma unroll
        for (int k = 0; k < kTile / 2 / kStepThreads; ++k) {
            const int j = threadIdx.x + k * kStepThreads;        // pair index inside the tile
            const int e0 = 2 * j;
            if (e0 < valid) {
                const bool two = (e0 + 1 < valid);
                double xd0 = 0, xd1 = 0, xr0 = 0, xr1 = 0;
                double2 u = make_double2(0.0, 0.0);
                int2 f = make_int2(0, 0), t = make_int2(0, 0);
                if (need_dec) { const V2 v = sxd[j]; xd0 = (double)v.x; xd1 = (double)v.y; }
                if (a.moments) { const V2 v = sxr[j]; xr0 = (double)v.x; xr1 = (double)v.y; }
                if (need_u_in) u = su[j];
                const int64_t p = p0 + e0;
                if (a.first_hit) { f.x = __ldg(a.first_hit + p); if (two) f.y = __ldg(a.first_hit + p + 1); }
                if (a.tau && need_u_in) { t.x = a.tau[p]; if (two) t.y = a.tau[p + 1]; }
                bool changed = path_step<D>(a, gam, xd0, xr0, u.x, t.x, f.x, acc);
                if (two) changed |= path_step<D>(a, gam, xd1, xr1, u.y, t.y, f.y, acc);
                // the state is written only where a path exercised (16-byte granularity): below maturity
                // most pairs are untouched, which removes most of the write traffic
                if (write_u && changed) {
                    if (two) {
                        if (a.l2_hints) st_hint(reinterpret_cast<double2*>(a.U + p), u, pol_keep);
                        else *reinterpret_cast<double2*>(a.U + p) = u;
                        if (a.tau) *reinterpret_cast<int2*>(a.tau + p) = t;
                    } else {
                        if (a.l2_hints) st_hint(a.U + p, u.x, pol_keep);
                        else a.U[p] = u.x;
                        if (a.tau) a.tau[p] = t.x;
                    }
                }
            }
        }
        __syncthreads();                 // every warp is done with stage s before it is refilled
    }
    block_reduce_store<NACC, kStepThreads, kAccStride>(acc, red, a.partials + (int64_t)blockIdx.x * kAccStride);
}

// ---------------------------------------------------------------------------------------------------------
// Solve kernel: <<<1, 128>>>.
//   phase 1 (do_reduce): sums[a] = sum over rows of partials[row][a], fixed order (lane-strided, then a
//                        shuffle tree) -> bitwise reproducible, no floating-point atomics anywhere.
//   phase 2 (do_solve):  thread 0 runs the k x k solve and stores gamma / diagnostics.
//   final_price:         price = sum(U) / P.
// Multi-GPU: phase 1, then an NCCL all-reduce of `sums`, then phase 2 as a second launch.
constexpr int kSolveThreads = 256;

template <int K>
__global__ void __launch_bounds__(kSolveThreads, 1) lsm_solve_kernel(const SolveArgs a) {
    constexpr int d = K - 1;
    constexpr int nacc = 3 * d + 1;
    __shared__ double part[kSolveThreads / 32][kAccStride];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    if (a.do_reduce) {
        // rows are 32 doubles: lane = accumulator, warp = row group; loads are coalesced and issued in batches
        // of 8 so the latency of the (L2-resident) partials is paid a handful of times, not once per row.
        // Summation order is fixed (row group, then rows ascending, then groups ascending): deterministic.
        double v = 0.0;
        int row = grp;
        for (; row + 7 * (kSolveThreads / 32) < a.n_rows; row += 8 * (kSolveThreads / 32)) {
            double t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) t[q] = a.partials[(int64_t)(row + q * (kSolveThreads / 32)) * kAccStride + lane];
#pragma unroll
            for (int q = 0; q < 8; ++q) v += t[q];
        }
        for (; row < a.n_rows; row += kSolveThreads / 32) v += a.partials[(int64_t)row * kAccStride + lane];
        part[grp][lane] = v;
        __syncthreads();
        if (threadIdx.x < nacc) {
            double tot = 0.0;
#pragma unroll
            for (int q = 0; q < kSolveThreads / 32; ++q) tot += part[q][threadIdx.x];
            a.sums[threadIdx.x] = tot;
            part[0][threadIdx.x] = tot;
        }
        __syncthreads();
    } else {
        if (threadIdx.x < nacc) part[0][threadIdx.x] = a.sums[threadIdx.x];
        __syncthreads();
    }
    if (a.peer.world > 1) {
        // fused all-reduce over peer memory: push this rank's sums into everybody's mailbox, then gather
        const int W = a.peer.world;
        const uint32_t seq = a.peer.seq;
        const int slot = (int)(seq % kPeerRing);
        for (int idx = threadIdx.x; idx < W * nacc; idx += kSolveThreads) {
            const int q = idx / nacc, i = idx - q * nacc;
            st_ll(a.peer.mailbox[q] + ((slot * W + a.peer.rank) * kAccStride + i), part[0][i], seq);
        }
        double tot = 0.0;
        if (threadIdx.x < nacc) {
            const uint4* mine = a.peer.mailbox[a.peer.rank] + (slot * W) * kAccStride + threadIdx.x;
            const uint64_t t0 = global_timer_ns();
            for (int q = 0; q < W; ++q) {                  // rank order: identical bits on every rank
                double v;
                int spins = 0;
                while (!ld_ll(mine + q * kAccStride, seq, v)) {
                    if (((++spins) & 1023) == 0 && global_timer_ns() - t0 > 4000000000ull) {   // 4 s: a peer is gone
                        *a.peer.err = 1;
                        v = 0.0;
                        break;
                    }
                }
                tot += v;
            }
        }
        __syncthreads();                                   // every thread has read part[0][*] for its pushes
        if (threadIdx.x < nacc) {
            part[0][threadIdx.x] = tot;
            a.sums[threadIdx.x] = tot;
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    if (a.final_price) {
        a.price[0] = part[0][2 * d] / a.spec.n_paths;
        return;
    }
    if (!a.do_solve) return;
    double h[2 * d + 1], g[K];
    h[0] = a.spec.n_paths;
#pragma unroll
    for (int m = 1; m <= 2 * d; ++m) h[m] = part[0][m - 1];
#pragma unroll
    for (int m = 0; m <= d; ++m) g[m] = part[0][2 * d + m];
    SolveResult res;
    lsm_solve_t<K>(a.spec, h, g, a.y_scale, a.mu_ref, a.sigma_ref, &res);
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
        a.gamma[i] = res.gamma[i];
        if (a.beta) a.beta[i] = res.beta[i];
        if (a.sv) a.sv[i] = res.sv[i];
    }
    if (a.mean_std) { a.mean_std[0] = res.mean_x; a.mean_std[1] = res.std_x; }
    if (a.rank) a.rank[0] = res.rank;
}

// ---------------------------------------------------------------------------------------------------------
// Continuation value of every path at one step from a stored polynomial (lazy replacement for the two [P]
// copies the reference appends per step, amc.py:164): out = clamp ? max(fit, 0) : fit.
template <typename XT>
__global__ void continuation_kernel(const XT* __restrict__ x, int64_t n, const double* __restrict__ gamma, int degree,
                                    double mu, double isg, int clamp, double* __restrict__ out) {
    double gam[kMaxK];
    for (int i = 0; i <= degree; ++i) gam[i] = gamma[i];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double z = ((double)x[p] - mu) * isg;
        double f = gam[degree];
        for (int m = degree - 1; m >= 0; --m) f = fma(f, z, gam[m]);
        // np.maximum propagates NaN; fmax would drop it
        out[p] = clamp ? ((f > 0.0 || f != f) ? f : 0.0) : f;
    }
}

__global__ void intrinsic_kernel(const double* __restrict__ S, int64_t n, double K, int is_put, double* __restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double v = is_put ? (K - S[p]) : (S[p] - K);
        out[p] = (v > 0.0 || v != v) ? v : 0.0;     // np.maximum(v, 0)
    }
}

// get_basis_polynomials (amc.py:98-106): out[p][j] = phi_j(x_p), three-term recurrences in x.
__global__ void basis_matrix_kernel(const double* __restrict__ X, int64_t n, int basis, int degree,
                                    double* __restrict__ out) {
    const int k = degree + 1;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double x = X[p];
        double* row = out + p * k;
        double pm2 = 0.0, pm1 = 1.0;
        row[0] = 1.0;
        for (int j = 1; j < k; ++j) {
            double v;
            switch (basis) {
                case kChebyshev: v = (j == 1) ? x : 2.0 * x * pm1 - pm2; break;
                case kLegendre: v = ((2.0 * j - 1.0) * x * pm1 - (j - 1.0) * pm2) / (double)j; break;
                case kLaguerre: v = ((2.0 * j - 1.0 - x) * pm1 - (j - 1.0) * pm2) / (double)j; break;
                default: v = pm1 * x;
            }
            row[j] = v;
            pm2 = pm1;
            pm1 = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// AMC_STEP_KERNEL=ldg selects the register-staged kernel (kept for A/B measurements); default is the TMA ring.
static bool use_tma_kernel() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AMC_STEP_KERNEL");
        v = (e && e[0] == 'l') ? 0 : 1;
    }
    return v == 1;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool pdl,
                             Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <typename XT, int D>
static cudaError_t launch_step_t(int grid, const StepArgs& a, cudaStream_t s, bool pdl) {
    if (use_tma_kernel()) {
        constexpr int smem = kStages * StageBytes<XT>::value;
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(lsm_step_tma_kernel<XT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            configured = true;
        }
        return launch_ex(lsm_step_tma_kernel<XT, D>, grid, kStepThreads, smem, s, pdl, a);
    } else {
        lsm_step_kernel<XT, D><<<grid, kStepThreads, 0, s>>>(a);
    }
    return cudaGetLastError();
}

template <typename XT>
static cudaError_t launch_step_d(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl) {
    switch (degree) {
        case 0: return launch_step_t<XT, 0>(grid, a, s, pdl);
        case 1: return launch_step_t<XT, 1>(grid, a, s, pdl);
        case 2: return launch_step_t<XT, 2>(grid, a, s, pdl);
        case 3: return launch_step_t<XT, 3>(grid, a, s, pdl);
        case 4: return launch_step_t<XT, 4>(grid, a, s, pdl);
        case 5: return launch_step_t<XT, 5>(grid, a, s, pdl);
        case 6: return launch_step_t<XT, 6>(grid, a, s, pdl);
        case 7: return launch_step_t<XT, 7>(grid, a, s, pdl);
        case 8: return launch_step_t<XT, 8>(grid, a, s, pdl);
        case 9: return launch_step_t<XT, 9>(grid, a, s, pdl);
        case 10: return launch_step_t<XT, 10>(grid, a, s, pdl);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_step(int dtype, int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl) {
    return dtype == 1 ? launch_step_d<float>(degree, grid, a, s, pdl) : launch_step_d<double>(degree, grid, a, s, pdl);
}

template <typename XT, int D>
static int occupancy_blocks() {
    int nb = 0;
    if (use_tma_kernel()) {
        constexpr int smem = kStages * StageBytes<XT>::value;
        cudaFuncSetAttribute(lsm_step_tma_kernel<XT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lsm_step_tma_kernel<XT, D>, kStepThreads, smem);
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lsm_step_kernel<XT, D>, kStepThreads, 0);
    }
    return nb;
}

template <typename XT>
static int occupancy_d(int degree) {
    switch (degree) {
        case 0: return occupancy_blocks<XT, 0>();
        case 1: return occupancy_blocks<XT, 1>();
        case 2: return occupancy_blocks<XT, 2>();
        case 3: return occupancy_blocks<XT, 3>();
        case 4: return occupancy_blocks<XT, 4>();
        case 5: return occupancy_blocks<XT, 5>();
        case 6: return occupancy_blocks<XT, 6>();
        case 7: return occupancy_blocks<XT, 7>();
        case 8: return occupancy_blocks<XT, 8>();
        case 9: return occupancy_blocks<XT, 9>();
        case 10: return occupancy_blocks<XT, 10>();
    }
    return 1;
}

// grid = SM count x resident blocks per SM: every block is co-resident, the grid-stride loop balances.
int step_grid_size(int dtype, int degree, int sm_count) {
    int nb = dtype == 1 ? occupancy_d<float>(degree) : occupancy_d<double>(degree);
    if (nb < 1) nb = 1;
    return sm_count * nb;
}

cudaError_t launch_solve(const SolveArgs& a, cudaStream_t s, bool pdl) {
    switch (a.spec.degree) {
        case 0: return launch_ex(lsm_solve_kernel<1>, 1, kSolveThreads, 0, s, pdl, a);
        case 1: return launch_ex(lsm_solve_kernel<2>, 1, kSolveThreads, 0, s, pdl, a);
        case 2: return launch_ex(lsm_solve_kernel<3>, 1, kSolveThreads, 0, s, pdl, a);
        case 3: return launch_ex(lsm_solve_kernel<4>, 1, kSolveThreads, 0, s, pdl, a);
        case 4: return launch_ex(lsm_solve_kernel<5>, 1, kSolveThreads, 0, s, pdl, a);
        case 5: return launch_ex(lsm_solve_kernel<6>, 1, kSolveThreads, 0, s, pdl, a);
        case 6: return launch_ex(lsm_solve_kernel<7>, 1, kSolveThreads, 0, s, pdl, a);
        case 7: return launch_ex(lsm_solve_kernel<8>, 1, kSolveThreads, 0, s, pdl, a);
        case 8: return launch_ex(lsm_solve_kernel<9>, 1, kSolveThreads, 0, s, pdl, a);
        case 9: return launch_ex(lsm_solve_kernel<10>, 1, kSolveThreads, 0, s, pdl, a);
        case 10: return launch_ex(lsm_solve_kernel<11>, 1, kSolveThreads, 0, s, pdl, a);
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

static inline int blocks_for(int64_t n, int threads, int cap) {
    int64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    return (int)(b > cap ? cap : b);
}

cudaError_t launch_continuation(int dtype, const void* x, int64_t n, const double* gamma_dev, int degree, double mu,
                                double isg, int clamp, double* out_dev, cudaStream_t s) {
    const int grid = blocks_for(n, 256, 148 * 8);
    if (dtype == 1)
        continuation_kernel<float><<<grid, 256, 0, s>>>((const float*)x, n, gamma_dev, degree, mu, isg, clamp, out_dev);
    else
        continuation_kernel<double><<<grid, 256, 0, s>>>((const double*)x, n, gamma_dev, degree, mu, isg, clamp, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_intrinsic(const double* S_dev, int64_t n, double K, int is_put, double* out_dev, cudaStream_t s) {
    intrinsic_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, s>>>(S_dev, n, K, is_put, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_basis_matrix(const double* X_dev, int64_t n, int basis, int degree, double* out_dev,
                                cudaStream_t s) {
    basis_matrix_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, s>>>(X_dev, n, basis, degree, out_dev);
    return cudaGetLastError();
}

}  // namespace amc
