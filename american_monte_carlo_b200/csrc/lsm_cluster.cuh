// Small path sets: the whole backward sweep inside ONE thread-block cluster (sm_100a).
//
// Below a few hundred thousand paths a step of the launch chain (lsm_step.cuh -> lsm_solve_kernel) costs 6-8 us, almost
// all of it the two grid hand-offs; the grid-wide persistent kernel (lsm_sweep.cuh) replaces them by hand-offs through L2
// and is no faster (profiles/r2_persistent_vs_chain.md).  What is left is hardware that synchronises without L2: a
// cluster of up to 16 CTAs on the SMs of one GPC.
//
//   * the per-path state U and two path columns live in the CTAs' shared memory for the whole sweep (CTA r owns the
//     paths [r * slice, (r + 1) * slice)); every column is fetched from global memory exactly once, by one 1-D bulk copy
//     per CTA (cp.async.bulk + mbarrier) issued a whole pass before it is needed: column t-2 replaces column t as soon as
//     the pass of step t is through its loop, and lands while that pass reduces and solves;
//   * pass t: decide(t) + moments(t-1) from shared memory (the arithmetic of the step kernel: path_step /
//     fast_path_step, 8/4/2 paths in flight per thread), block reduction by recursive halving into this CTA's row, which
//     the CTA PUSHES into the shared memory of every CTA of the cluster with asynchronous remote stores that complete the
//     transaction count of the receiver's mbarrier (st.async ... mbarrier::complete_tx::bytes); a CTA waits on its OWN
//     mbarrier only -- no cluster-wide barrier inside the sweep -- then adds the rows pairwise in rank order and runs the
//     k x k solve itself (same inputs, same code, same bits: no broadcast); rows are double-buffered by pass parity;
//   * the last CTA (shortest slice) copies the regression diagnostics of the step and, after the last pass, the price to
//     global memory; every CTA writes its slice of the state back once at the end.
//
// No spin loops, no global flags: the only waits are the CTA's own mbarriers (column copies, rows) and the hardware cluster
// barrier at the start and the end of the kernel.
// Capacity is what 16 x ~215 KB of shared memory hold (2 columns + state per path): ~146k paths f64/f64, ~290k f32/f32;
// the default policy (api.cu) stops at 147456 paths, where the launch chain's 148 SMs catch up with the cluster's 16
// (profiles/r2_cluster_vs_chain.md), and at degree 5 (beyond, the warp-cooperative solve is used and the chain is faster).
#pragma once
#include <mutex>
#include <type_traits>

#include "kernels.h"
#include "lsm_solve_block.cuh"
#include "lsm_step.cuh"

namespace amc {

constexpr int kClusterThreads = kSolveThreads;      // solve_block() is written for this block size
static_assert(kClusterThreads == 256, "the block reduction of lsm_cluster_kernel adds 8 warps pairwise");
constexpr int kClusterMaxCtas = 16;

__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_cta_count() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster; orders shared-memory writes before it against reads (local or remote) after it
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// asynchronous store of 8 bytes into the shared memory of CTA `cta` of the cluster (at the address `local` has in this
// CTA) that completes 8 bytes of the pending transaction of THAT CTA's mbarrier `local_bar`: the receiver waits on its
// own barrier -- no cluster-wide barrier, no fence
__device__ __forceinline__ void st_async_remote_f64(double* local, uint64_t* local_bar, uint32_t cta, double v) {
    uint32_t remote, remote_bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr(local)), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_bar) : "r"(smem_addr(local_bar)), "r"(cta));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}

// Sum N per-lane values over the 32 lanes of a warp by recursive halving: at every stage a lane hands half of its vector
// to the partner lane and adds what it receives to the half it keeps -- NP - 1 + log2(32 / NP) shuffles for NP = N rounded
// up to a power of two, instead of 5 N butterflies.  On return v[0] of lane l is the warp total of accumulator
// warp_sum_slot<N>(l) (slots >= N are padding); every accumulator is held by 32 / NP lanes.  Fixed order: deterministic.
template <int N> struct WarpSumPad { static constexpr int value = N <= 1 ? 1 : (N <= 2 ? 2 : (N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 16 ? 16 : 32)))); };
template <int N>
__device__ __forceinline__ int warp_sum_slot(int lane) {
    constexpr int NP = WarpSumPad<N>::value;
    int slot = 0;
    // stage k uses lane bit 16 >> k and decides bit (NP / 2) >> k of the slot
#pragma unroll
    for (int b = 16, h = NP / 2; h >= 1; b >>= 1, h >>= 1) slot += (lane & b) ? h : 0;
    return slot;
}
template <int N>
__device__ __forceinline__ double warp_transpose_sum(const double (&acc)[N]) {
    constexpr int NP = WarpSumPad<N>::value;
    const int lane = threadIdx.x & 31;
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) v[i] = (i < N) ? acc[i] : 0.0;
    int b = 16;
#pragma unroll
    for (int len = NP; len > 1; len >>= 1, b >>= 1) {
        const bool up = (lane & b) != 0;
#pragma unroll
        for (int i = 0; i < len / 2; ++i) {
            const double send = up ? v[i] : v[i + len / 2];
            const double keep = up ? v[i + len / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, b);
        }
    }
    // the lane bits not used for splitting: plain butterflies on the one value left
#pragma unroll
    for (; b >= 1; b >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], b);
    return v[0];
}

// paths per CTA: an equal share, padded to 32 elements like the columns themselves (bulk copies are 128-byte granular)
__host__ __device__ inline int64_t cluster_slice(int64_t n_paths, int n_ctas) {
    return ((n_paths + n_ctas - 1) / n_ctas + 31) / 32 * 32;
}
template <typename XT, typename UT>
constexpr size_t cluster_bytes_per_path() { return 2 * sizeof(XT) + sizeof(UT); }

template <typename XT, typename UT, int D>
__global__ void __launch_bounds__(kClusterThreads, 1) lsm_cluster_kernel(const __grid_constant__ SweepArgs a) {
    constexpr int K = D + 1;
    constexpr int NACC = 3 * D + 1;
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ double red[(kClusterThreads / 32) * WarpSumPad<NACC>::value];
    // rows[p & 1][q] = CTA q's partial sums of pass p, PUSHED here by CTA q with asynchronous remote stores that complete
    // the transaction count of rowbar[p & 1] (every CTA holds all rows and waits on its own barrier only; a CTA can push
    // the rows of pass p + 2 only after every CTA has read those of pass p -- it needs everybody's rows of pass p + 1 first)
    __shared__ __align__(16) double rows[2][kClusterMaxCtas][kAccStride];
    __shared__ uint64_t rowbar[2];
    __shared__ double sums_sh[kAccStride];
    __shared__ double o_gamma[kMaxK], o_beta[kMaxK], o_sv[kMaxK], o_ms[4], o_price[1];
    __shared__ int o_rank[1];
    __shared__ uint64_t full[2];                              // column t lands in buffer t & 1

    const uint32_t cta = cluster_cta_rank(), n_ctas = cluster_cta_count();
    const int64_t slice = cluster_slice(a.n_paths, (int)n_ctas);
    const int64_t p_lo = (int64_t)cta * slice;
    int64_t left = a.n_paths - p_lo;
    const int cnt = (int)(left < 0 ? 0 : (left > slice ? slice : left));
    const uint32_t col_bytes = (uint32_t)((cnt + 31) / 32 * 32) * (uint32_t)sizeof(XT);

    XT* const xs0 = reinterpret_cast<XT*>(dyn);
    XT* const xs1 = xs0 + slice;
    UT* const us = reinterpret_cast<UT*>(xs1 + slice);
    const int n = a.n_steps;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&rowbar[0], 1);
        mbar_init(&rowbar[1], 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 2 * kClusterMaxCtas * kAccStride; i += kClusterThreads) (&rows[0][0][0])[i] = 0.0;
    cluster_barrier();                           // every CTA of the cluster is running: its shared memory may be written
    // thread 0: fetch column t into its buffer
    auto fetch = [&](int t) {
        if (cnt == 0) return;
        const XT* src = reinterpret_cast<const XT*>(static_cast<const char*>(a.S) + (size_t)t * (size_t)a.ld * sizeof(XT)) + p_lo;
        mbar_expect_tx(&full[t & 1], col_bytes);
        tma_load_1d((t & 1) ? xs1 : xs0, src, col_bytes, &full[t & 1]);
    };
    if (threadIdx.x == 0) {
        fetch(n);
        if (a.n_passes > 1 && n >= 1) fetch(n - 1);
    }
    FastConsts fc;
    fc.sgn = a.is_put ? -1.0 : 1.0;
    fc.sgnK = a.is_put ? a.K : -a.K;
    // per-column constants: the entry of column t-1 is the next pass's entry of column t, the one after is fetched a
    // whole pass ahead (its global-memory latency is off the pass-to-pass critical path)
    SweepTab td = a.tab[n];
    SweepTab tr = (a.n_passes > 1 && n >= 1) ? a.tab[n - 1] : td;

    // AMC_CLUSTER_TRACE (api.cu): CTA 0 / thread 0 leaves its SM clock at 8 points of every pass
    long long* const trace = (cta == 0 && threadIdx.x == 0) ? reinterpret_cast<long long*>(a.sync) : nullptr;

    for (int p = 0; p < a.n_passes; ++p) {
        const int t = n - p;
        const int mode = (p == 0) ? kMaturity : (a.american ? kDecide : kObserve);
        const bool moments = a.n_passes > 1 && t > 0;
        const bool final_pass = (p == a.n_passes - 1);
        if (trace) trace[p * 8 + 0] = clock64();
        const SweepTab tr_next = (a.n_passes > 1 && t >= 2) ? a.tab[t - 2] : tr;
        SolverTab tb;
        tb.y_scale = 1.0; tb.mu = 0.0; tb.sigma = 1.0; tb.pad = 0.0;
        if (!final_pass) tb = a.solve_tab[t - 1];
        double gam[D + 1];
#pragma unroll
        for (int i = 0; i <= D; ++i) gam[i] = (mode == kDecide) ? o_gamma[i] : 0.0;     // left by the solve of pass p-1

        StepArgs sa = {};                        // the launch-uniform view path_step() expects
        sa.t_dec = t;
        sa.mode = mode;
        sa.moments = moments ? 1 : 0;
        sa.is_put = a.is_put;
        sa.K = a.K;
        sa.disc_dec = td.disc;
        sa.mu_dec = td.mu; sa.isg_dec = td.isg;
        sa.mu_reg = tr.mu; sa.isg_reg = tr.isg;
        fc.da = td.isg; fc.db = -td.mu * td.isg;
        fc.ra = tr.isg; fc.rb = -tr.mu * tr.isg;
        fc.disc = td.disc;
        const bool fast_ok = (mode == kDecide) && moments && !a.first_hit;

        const XT* const xd = (t & 1) ? xs1 : xs0;
        const XT* const xr = (t & 1) ? xs0 : xs1;
        if (cnt > 0) {
            // column t is the ((n - t) / 2)-th copy into its buffer (waiting twice for the same copy is fine: no later
            // copy into that buffer has been issued yet)
            mbar_wait(&full[t & 1], (uint32_t)((n - t) >> 1) & 1u);
            if (moments) mbar_wait(&full[(t - 1) & 1], (uint32_t)((n - t + 1) >> 1) & 1u);
        }
        if (trace) trace[p * 8 + 1] = clock64();
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
        // Several paths per thread and iteration, loaded first and then stepped: with 8 warps per SM the loop is bound by the
        // latency of one path's dependent FP64 chain unless several chains are in flight per thread.
        int j = threadIdx.x;
        auto fast_chunks = [&](auto ilp_c) {
            constexpr int kIlp = decltype(ilp_c)::value;
            for (; j + (kIlp - 1) * kClusterThreads < cnt; j += kIlp * kClusterThreads) {
                double x_dec[kIlp], x_reg[kIlp], u[kIlp];
                bool changed[kIlp];
#pragma unroll
                for (int k = 0; k < kIlp; ++k) {
                    x_dec[k] = (double)xd[j + k * kClusterThreads];
                    x_reg[k] = (double)xr[j + k * kClusterThreads];
                    u[k] = (double)us[j + k * kClusterThreads];
                }
#pragma unroll
                for (int k = 0; k < kIlp; ++k) changed[k] = fast_path_step<D>(fc, gam, x_dec[k], x_reg[k], u[k], true, acc);
#pragma unroll
                for (int k = 0; k < kIlp; ++k) {
                    if (changed[k]) {
                        us[j + k * kClusterThreads] = (UT)u[k];
                        if (a.tau) a.tau[p_lo + j + k * kClusterThreads] = t;
                    }
                }
            }
        };
        if (fast_ok) {
            if (D <= 3) fast_chunks(std::integral_constant<int, 8>{});
            fast_chunks(std::integral_constant<int, 4>{});
            fast_chunks(std::integral_constant<int, 2>{});
        }
        for (; j < cnt; j += kClusterThreads) {
            const double x_dec = (double)xd[j];
            const double x_reg = moments ? (double)xr[j] : 0.0;
            double u = (p != 0) ? (double)us[j] : 0.0;
            bool changed;
            int tau_j = t;
            if (fast_ok) {
                changed = fast_path_step<D>(fc, gam, x_dec, x_reg, u, true, acc);
            } else {
                const int fh = a.first_hit ? __ldg(a.first_hit + p_lo + j) : 0;
                changed = path_step<D>(sa, gam, x_dec, x_reg, u, tau_j, fh, acc);
            }
            if (changed) {
                us[j] = (UT)u;
                if (a.tau) a.tau[p_lo + j] = t;
            }
        }
        __syncthreads();                         // everybody is through with column t: its buffer takes column t-2
        if (threadIdx.x == 0 && a.n_passes > 1 && t >= 2) fetch(t - 2);
        if (trace) trace[p * 8 + 2] = clock64();

        {
            // block reduction into this CTA's row: warp totals by recursive halving, then the 8 warps' values pairwise
            constexpr int NP = WarpSumPad<NACC>::value;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const double wtot = warp_transpose_sum<NACC>(acc);
            const int slot = warp_sum_slot<NACC>(lane);
            if ((lane & (32 / NP - 1)) == 0) red[warp * NP + slot] = wtot;      // one of the lanes holding this slot
            __syncthreads();
            if (threadIdx.x < kAccStride) {
                double r = 0.0;
                if (threadIdx.x < NACC) {
                    const double* q = red + threadIdx.x;
                    r = ((q[0] + q[NP]) + (q[2 * NP] + q[3 * NP])) + ((q[4 * NP] + q[5 * NP]) + (q[6 * NP] + q[7 * NP]));
                }
                if (threadIdx.x == 0) mbar_expect_tx(&rowbar[p & 1], n_ctas * (uint32_t)NACC * 8u);    // what this CTA receives
                if (threadIdx.x < NACC) {
#pragma unroll
                    for (int q = 0; q < kClusterMaxCtas; ++q)
                        if (q < (int)n_ctas) st_async_remote_f64(&rows[p & 1][cta][threadIdx.x], &rowbar[p & 1], (uint32_t)q, r);
                }
            }
        }
        if (trace) trace[p * 8 + 3] = clock64();

        if (threadIdx.x < kAccStride) {
            mbar_wait(&rowbar[p & 1], (uint32_t)(p >> 1) & 1u);     // every CTA's row of this pass has landed here
            if (trace) trace[p * 8 + 4] = clock64();
            double part[kClusterMaxCtas];
#pragma unroll
            for (int q = 0; q < kClusterMaxCtas; ++q)
                part[q] = (q < (int)n_ctas) ? rows[p & 1][q][threadIdx.x] : 0.0;
            // pairwise, in an order that depends on nothing but the ranks: the same bits in every CTA
#pragma unroll
            for (int w = 1; w < kClusterMaxCtas; w *= 2) {
#pragma unroll
                for (int q = 0; q + w < kClusterMaxCtas; q += 2 * w) part[q] += part[q + w];
            }
            sums_sh[threadIdx.x] = part[0];
            if (threadIdx.x < kMaxK) { o_gamma[threadIdx.x] = 0.0; o_beta[threadIdx.x] = 0.0; o_sv[threadIdx.x] = 0.0; }
            if (threadIdx.x < 4) o_ms[threadIdx.x] = 0.0;
            if (threadIdx.x == 0) o_rank[0] = 0;
        }
        __syncthreads();
        if (trace) trace[p * 8 + 5] = clock64();

        const int row = final_pass ? 0 : t - 1;
        SolveArgs sv = {};
        sv.sums = sums_sh;
        sv.do_reduce = 0;
        sv.do_solve = final_pass ? 0 : 1;
        sv.final_price = final_pass ? 1 : 0;
        sv.spec = a.solve.spec;
        sv.y_scale = tb.y_scale; sv.mu_ref = tb.mu; sv.sigma_ref = tb.sigma;
        sv.gamma = o_gamma; sv.beta = o_beta; sv.sv = o_sv; sv.mean_std = o_ms; sv.rank = o_rank; sv.price = o_price;
        sv.n_batch = 1;
        // (the scalar routine inlined with its matrices in registers, as in the dedicated solve kernel)
        solve_block<K, false>(sv, sv, 0u, nullptr);          // ends with a block barrier: o_* are complete
        if (trace) trace[p * 8 + 6] = clock64();

        if (cta == n_ctas - 1) {                  // the CTA with the shortest slice keeps the books
            if (final_pass) {
                if (threadIdx.x == 0) {
                    a.solve.price[0] = o_price[0];
                    a.solve.sums[2 * D] = sums_sh[2 * D];
                }
            } else if (threadIdx.x < kMaxK) {
                const size_t o = (size_t)row * kMaxK + threadIdx.x;
                a.solve.gamma[o] = o_gamma[threadIdx.x];
                if (a.solve.beta) a.solve.beta[o] = o_beta[threadIdx.x];
                if (a.solve.sv) a.solve.sv[o] = o_sv[threadIdx.x];
                if (threadIdx.x < 3 && a.solve.mean_std) a.solve.mean_std[(size_t)row * 3 + threadIdx.x] = o_ms[threadIdx.x];
                if (threadIdx.x == 0 && a.solve.rank) a.solve.rank[row] = o_rank[0];
            }
        }
        td = tr;
        tr = tr_next;
        if (trace) trace[p * 8 + 7] = clock64();
    }
    // the state goes back to global memory once (cashflows are an output of the sweep)
    UT* const Ug = static_cast<UT*>(a.U) + p_lo;
    for (int j = threadIdx.x; j < cnt; j += kClusterThreads) Ug[j] = us[j];
    cluster_barrier();                           // nobody leaves while a neighbour may still be reading its rows
}

// ---------------------------------------------------------------------------------------------------------
// host side: capacity query and launch
struct ClusterPlan {
    int n_ctas = 0;             // 0: this kernel cannot run here (no cluster of >= 2 CTAs with the shared memory it needs)
    size_t dyn_max = 0;         // dynamic shared memory per CTA the kernel may use
    int64_t max_paths = 0;
};

template <typename XT, typename UT, int D>
static ClusterPlan make_cluster_plan(int dev) {
    ClusterPlan plan;
    auto kernel = lsm_cluster_kernel<XT, UT, D>;
    cudaFuncAttributes fa;
    int optin = 0;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess ||
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
        cudaGetLastError();
        return plan;
    }
    const int64_t dyn = ((int64_t)optin - (int64_t)fa.sharedSizeBytes - 1024) / 128 * 128;
    if (dyn < 16384) return plan;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess ||
        cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return plan;
    }
    for (int nc = kClusterMaxCtas; nc >= 2; nc /= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nc);
        cfg.blockDim = dim3(kClusterThreads);
        cfg.dynamicSmemBytes = (size_t)dyn;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = nc;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, kernel, &cfg) == cudaSuccess && n_clusters >= 1) {
            plan.n_ctas = nc;
            plan.dyn_max = (size_t)dyn;
            const int64_t slice_max = (dyn / (int64_t)cluster_bytes_per_path<XT, UT>()) / 32 * 32;
            plan.max_paths = slice_max * nc;
            break;
        }
        cudaGetLastError();
    }
    return plan;
}

// the plan of the CURRENT device (function attributes are per device), made on first use; safe from several host threads
template <typename XT, typename UT, int D>
static const ClusterPlan& cluster_plan_t() {
    constexpr int kMaxDevices = 64;
    static std::mutex mu;
    static ClusterPlan plans[kMaxDevices];
    static bool made[kMaxDevices] = {};
    static const ClusterPlan none;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        cudaGetLastError();
        return none;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (!made[dev]) {
        plans[dev] = make_cluster_plan<XT, UT, D>(dev);
        made[dev] = true;
    }
    return plans[dev];
}

template <typename XT, typename UT, int D>
static cudaError_t launch_cluster_t(const SweepArgs& a, cudaStream_t s) {
    const ClusterPlan& plan = cluster_plan_t<XT, UT, D>();
    if (plan.n_ctas == 0 || a.n_paths > plan.max_paths || a.n_paths < 1) return cudaErrorInvalidConfiguration;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.n_ctas);
    cfg.blockDim = dim3(kClusterThreads);
    cfg.dynamicSmemBytes = (size_t)cluster_slice(a.n_paths, plan.n_ctas) * cluster_bytes_per_path<XT, UT>();
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.n_ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, lsm_cluster_kernel<XT, UT, D>, a);
}

// Degrees 0..5 only: there the register-resident scalar routine is the solve (as in the dedicated solve kernel).  From
// degree 6 on the warp-cooperative routine is, and measured on B200 the chain is faster at every path count
// (profiles/r2_cluster_vs_chain.md), so those sets stay on it.
constexpr int kClusterMaxDegree = 5;

template <typename XT, typename UT>
static cudaError_t launch_cluster_d(int degree, const SweepArgs& a, cudaStream_t s) {
    if (degree < 0 || degree > kClusterMaxDegree) return cudaErrorInvalidValue;
    switch (degree) {
        case 0: return launch_cluster_t<XT, UT, 0>(a, s);
        case 1: return launch_cluster_t<XT, UT, 1>(a, s);
        case 2: return launch_cluster_t<XT, UT, 2>(a, s);
        case 3: return launch_cluster_t<XT, UT, 3>(a, s);
        case 4: return launch_cluster_t<XT, UT, 4>(a, s);
        case 5: return launch_cluster_t<XT, UT, 5>(a, s);
    }
    return cudaErrorInvalidValue;
}

template <typename XT, typename UT>
static int64_t cluster_capacity_d(int degree) {
    if (degree < 0 || degree > kClusterMaxDegree) return 0;
    switch (degree) {
        case 0: return cluster_plan_t<XT, UT, 0>().max_paths;
        case 1: return cluster_plan_t<XT, UT, 1>().max_paths;
        case 2: return cluster_plan_t<XT, UT, 2>().max_paths;
        case 3: return cluster_plan_t<XT, UT, 3>().max_paths;
        case 4: return cluster_plan_t<XT, UT, 4>().max_paths;
        case 5: return cluster_plan_t<XT, UT, 5>().max_paths;
    }
    return 0;
}

}  // namespace amc
