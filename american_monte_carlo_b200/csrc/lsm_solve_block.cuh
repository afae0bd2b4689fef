// The per-step solve as executed by ONE thread block (kSolveThreads threads): fixed-order reduction of the workers'
// partial rows, the multi-GPU exchange over peer memory, then the k x k solve (lsm_solve_warp.cuh / lsm_solve.h).
// Used by the per-launch solve kernel (contract batches, array ops) and, inside the persistent sweep kernel, by whichever
// worker block finishes a pass last.
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "lsm_solve_warp.cuh"

namespace amc {

constexpr int kSolveThreads = 256;

// One solve step by one block of kSolveThreads threads (shared by the per-launch kernel and the persistent solver).
// `seq` != 0: exchange the reduced sums with the peers under that sequence number (multi-GPU); `sync` (may be null) is
// the sweep's sync block, whose abort word is raised when a peer does not show up.
// The scalar routine keeps its k x k matrices in registers when it has a kernel to itself; inside the persistent sweep
// kernel it is compiled out of line under that kernel's register budget, so only the steps that need it (degenerate or
// rank-truncated columns, SVD diagnostics) pay for the spills and the streaming loop keeps its occupancy.
template <int K>
__device__ __noinline__ void solve_scalar_outofline(const SolveArgs& a, const double* part0) {
    constexpr int d = K - 1;
    double h[2 * d + 1], g[K];
    h[0] = a.spec.n_paths;
#pragma unroll
    for (int m = 1; m <= 2 * d; ++m) h[m] = part0[m - 1];
#pragma unroll
    for (int m = 0; m <= d; ++m) g[m] = part0[2 * d + m];
    SolveResult res;
    lsm_solve_t<K>(a.spec, h, g, a.y_scale, a.mu_ref, a.sigma_ref, &res);
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
        a.gamma[i] = res.gamma[i];
        if (a.beta) a.beta[i] = res.beta[i];
        if (a.sv) a.sv[i] = res.sv[i];
    }
    if (a.mean_std) { a.mean_std[0] = res.mean_x; a.mean_std[1] = res.std_x; a.mean_std[2] = res.pivot_loss; }
    if (a.rank) a.rank[0] = res.rank;
}

template <int K>
__device__ __forceinline__ void solve_scalar_inline(const SolveArgs& a, const double* part0) {
    constexpr int d = K - 1;
    double h[2 * d + 1], g[K];
    h[0] = a.spec.n_paths;
#pragma unroll
    for (int m = 1; m <= 2 * d; ++m) h[m] = part0[m - 1];
#pragma unroll
    for (int m = 0; m <= d; ++m) g[m] = part0[2 * d + m];
    SolveResult res;
    lsm_solve_t<K>(a.spec, h, g, a.y_scale, a.mu_ref, a.sigma_ref, &res);
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
        a.gamma[i] = res.gamma[i];
        if (a.beta) a.beta[i] = res.beta[i];
        if (a.sv) a.sv[i] = res.sv[i];
    }
    if (a.mean_std) { a.mean_std[0] = res.mean_x; a.mean_std[1] = res.std_x; a.mean_std[2] = res.pivot_loss; }
    if (a.rank) a.rank[0] = res.rank;
}

// OUT_OF_LINE: the scalar routine as a call (persistent sweep kernel) or inlined with its matrices in registers (the
// dedicated solve kernel, which has a block to itself)
template <int K, bool OUT_OF_LINE>
__device__ __forceinline__ void solve_block(const SolveArgs& a, const SolveArgs& a_in, uint32_t seq, uint32_t* sync) {
    constexpr int d = K - 1;
    constexpr int nacc = 3 * d + 1;
    __shared__ double part[kSolveThreads / 32][kAccStride];
    __shared__ SolveShared<K> solve_sh;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    if (a.do_reduce) {
        // A partial row is 32 doubles of which the first nacc are used: LPR lanes read one row, so a warp covers
        // 32 / LPR rows per load.  Every thread issues 16 independent (predicated) loads per round: the latency of the
        // L2-resident partials is paid ceil(rows / (16 * slots)) times -- 3 rounds for a full grid -- not once per row.
        // Summation order is fixed (slot, then rows ascending, then slots ascending): deterministic.  Loads bypass L1
        // (ld.cg): in the persistent sweep the rows are rewritten by other SMs between two reads of this block.
        constexpr int LPR = nacc <= 8 ? 8 : (nacc <= 16 ? 16 : 32);
        constexpr int RPW = 32 / LPR;
        constexpr int NSLOT = (kSolveThreads / 32) * RPW;
        const int col = lane % LPR;
        const int slot = grp * RPW + lane / LPR;
        // contract batches: the power sums (columns < 2d) are taken from contract 0's rows (see lsm_step.cuh)
        const double* src = (col < 2 * d) ? a_in.partials : a.partials;
        double v = 0.0;
        for (int row0 = slot; row0 < a.n_rows; row0 += 16 * NSLOT) {
            double t[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int r = row0 + q * NSLOT;
                t[q] = (r < a.n_rows) ? __ldcg(src + (int64_t)r * kAccStride + col) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) v += t[q];
        }
        double* flat = &part[0][0];                       // [NSLOT][LPR] = 256 doubles
        flat[slot * LPR + col] = v;
        __syncthreads();
        double tot = 0.0;
        if (threadIdx.x < nacc) {
#pragma unroll
            for (int q = 0; q < NSLOT; ++q) tot += flat[q * LPR + threadIdx.x];
        }
        __syncthreads();
        if (threadIdx.x < nacc) {
            a.sums[threadIdx.x] = tot;
            part[0][threadIdx.x] = tot;
        }
        __syncthreads();
    } else {
        if (threadIdx.x < nacc) part[0][threadIdx.x] = a.sums[threadIdx.x];
        __syncthreads();
    }
    if (seq != 0u && a.peer.world > 1) {
        // fused all-reduce over peer memory: push this rank's sums into everybody's mailbox, then gather
        const int W = a.peer.world;
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
                        if (sync) atomicExch(sync + 32, 1u);       // kSyncAbort: release the workers as well
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
    if (a.final_price) {
        if (threadIdx.x == 0) a.price[0] = part[0][2 * d] / a.spec.n_paths;
    } else if (a.do_solve && threadIdx.x < 32) {
        // warp 0: cooperative solve of the certified full-rank case; everything else (degenerate column, rank
        // truncation, SVD diagnostics) falls through to the scalar routine on thread 0
        const bool solved = lsm_solve_warp<K>(a.spec, &part[0][0], &part[0][2 * d], a.y_scale, a.mu_ref, a.sigma_ref, solve_sh,
                                              a.gamma, a.beta, a.sv, a.mean_std, a.rank);
        if (!solved && threadIdx.x == 0) {
            if (OUT_OF_LINE) solve_scalar_outofline<K>(a, &part[0][0]);
            else solve_scalar_inline<K>(a, &part[0][0]);
        }
    }
    __syncthreads();
}

}  // namespace amc
