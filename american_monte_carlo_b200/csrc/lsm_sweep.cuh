// Persistent backward sweep for sm_100a: ONE launch walks all n+1 time steps.
//
// The per-step launch chain of round 1 (step kernel -> solve kernel -> step kernel ..., 2(n+1) launches) paid two grid
// hand-offs, a pipeline fill and a pipeline drain per time step (~6 us fixed).  Here ONE cooperative launch (every block
// resident: SM count x occupancy) keeps the step kernel's blocks alive for the whole sweep:
//
//   block, pass p (t = n - p):          [columns of pass p already in flight]
//       wait  sync.published >= p            (the continuation polynomial of step t exists; p = 0: nothing to wait for)
//       decide(t) + moments(t-1) over its own tiles, TMA ring as before
//       partial row -> global, fence, ticket[p] += 1
//       issue the COLUMN copies of pass p+1's first tiles (immutable data: no dependency on the solve)
//       the block that drew the LAST ticket of the pass: reduce all rows in fixed order, (multi-GPU: peer-memory
//       exchange), solve, store gamma[t-1], fence, sync.published = p + 1     (last pass: price = sum(U) / P)
//
// Tile ownership is static (block b owns tiles b, b + G, ... in every pass), so the per-path state written in pass p is
// re-read by the same block in pass p+1: no inter-block hazard on U; the only global dependency per step is the
// polynomial, which travels through L2 (release/acquire), ~1 us instead of two kernel boundaries.  The solve runs
// inside this kernel's register budget: the warp-cooperative routine (matrices in shared memory) covers the certified
// full-rank steps; the scalar routine behind it (degenerate / rank-truncated steps) spills, which only those steps pay.
// Every spin loop has a wall-clock limit (kSyncAbort): a missing peer turns into an error code, never a hang.
//
// LEAN = true is the path-free sweep: no path matrix exists; the state per path is its
// fixed-point log2-price L_t (int32, gbm_quad.cuh) and the step's increments are regenerated from the Philox counters,
// L_{t-1} = L_t - q_t -- the exact reverse of the forward sum, so the prices (and decisions) equal the stored mode's.
#pragma once
#include "gbm_quad.cuh"
#include "kernels.h"
#include "lsm_solve_block.cuh"
#include "lsm_step.cuh"

namespace amc {

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// spin until *word >= want (or the sweep is aborted / the limit expires -> abort); returns false on abort
__device__ __forceinline__ bool spin_until_at_least(const uint32_t* word, uint32_t want, uint32_t* sync) {
    if (ld_acquire_u32(word) >= want) return true;
    const uint64_t t0 = global_timer_ns();
    for (uint32_t spins = 1;; ++spins) {
        if (ld_acquire_u32(word) >= want) return true;
        if ((spins & 255u) == 0u) {
            if (ld_acquire_u32(sync + kSyncAbort) != 0u) return false;
            if (global_timer_ns() - t0 > kSpinLimitNs) {
                atomicExch(sync + kSyncAbort, 1u);
                return false;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The solve of one pass, out of line: its register needs (row reduction, exchange, k x k solve) must not shape the
// register allocation of the streaming loop; the call happens once per pass in one block.
template <int K>
__device__ __noinline__ void sweep_solve_pass(const SweepArgs& a, int p, int t) {
    const bool final_pass = (p == a.n_passes - 1);
    SolveArgs sv = a.solve;
    sv.do_reduce = 1;
    sv.do_solve = final_pass ? 0 : 1;
    sv.final_price = final_pass ? 1 : 0;
    const int row = final_pass ? 0 : t - 1;
    const SolverTab tb = a.solve_tab[row];
    sv.y_scale = final_pass ? 1.0 : tb.y_scale;
    sv.mu_ref = final_pass ? 0.0 : tb.mu;
    sv.sigma_ref = final_pass ? 1.0 : tb.sigma;
    sv.gamma = a.solve.gamma + (size_t)row * kMaxK;
    sv.beta = a.solve.beta ? a.solve.beta + (size_t)row * kMaxK : nullptr;
    sv.sv = a.solve.sv ? a.solve.sv + (size_t)row * kMaxK : nullptr;
    sv.mean_std = a.solve.mean_std ? a.solve.mean_std + (size_t)row * 3 : nullptr;
    sv.rank = a.solve.rank ? a.solve.rank + row : nullptr;
    const uint32_t seq = (a.solve.peer.world > 1) ? a.seq_base + (uint32_t)p + 1u : 0u;
    solve_block<K, true>(sv, sv, seq, a.sync);
    if (threadIdx.x == 0) {
        __threadfence();
        st_release_u32(a.sync + kSyncPublished, (uint32_t)p + 1u);
    }
}

template <typename XT, typename UT, bool LEAN>
struct SweepStage {
    // stored: [x_dec tile][x_reg tile][U tile];  lean: [L tile (int32)][U tile]
    static constexpr int kColBytes = LEAN ? kTile * 4 : kTile * (int)sizeof(XT);
    static constexpr int kUOff = LEAN ? kColBytes : 2 * kColBytes;
    static constexpr int value = kUOff + kTile * (int)sizeof(UT);
};

// Resident blocks per SM the streaming loop is compiled for (the register cap that goes with it also bounds the
// out-of-line scalar solve): float columns 4 blocks up to degree 3 (<= 64 registers, as the per-launch kernel of round
// 1), 3 up to degree 5, else 2; double columns 2 (their 96 KB ring allows no more).
template <typename XT, int D, bool LEAN>
struct SweepMinBlocks {
#ifdef AMC_LEAN_BLOCKS
    static constexpr int kLean = (D <= 3) ? AMC_LEAN_BLOCKS : 2;
#else
    static constexpr int kLean = 2;
#endif
    static constexpr int value = LEAN ? kLean : (sizeof(XT) == 4 ? (D <= 3 ? 4 : (D <= 5 ? 3 : 2)) : 2);
};

// Geometry of one block's share of the sweep and thread 0's copy engine driver (all trivially inlined).
template <typename XT, typename UT, bool LEAN>
struct SweepRing {
    static constexpr int kStage = SweepStage<XT, UT, LEAN>::value;
    static constexpr int kUOff = SweepStage<XT, UT, LEAN>::kUOff;
    const SweepArgs* a;
    unsigned char* ring;
    uint64_t* full;
    int my_tiles;

    __device__ __forceinline__ int pass_t(int p) const { return a->n_steps - p; }
    __device__ __forceinline__ bool need_dec(int p) const { return LEAN ? true : (p == 0 || a->american != 0); }
    __device__ __forceinline__ bool moments(int p) const { return a->n_passes > 1 && pass_t(p) > 0; }
    // path offset of this block's i-th tile in pass p (static ownership: tiles blockIdx.x + k * gridDim.x; the direction
    // alternates from pass to pass so that the tail of pass p, still in L2, is the head of pass p+1)
    __device__ __forceinline__ int64_t tile_offset(int p, int i) const {
        const bool rev = a->reverse && (p & 1);
        return ((int64_t)blockIdx.x + (int64_t)(rev ? (my_tiles - 1 - i) : i) * gridDim.x) * kTile;
    }
    // thread 0: copies of item (p, i) into its ring stage.  parts bit 0: arm the barrier with the item's total byte count
    // and copy what does not depend on the previous pass (the path columns); bit 1: the state (U, and L when LEAN).
    __device__ __forceinline__ void issue(int p, int i, uint32_t item, int parts) const {
        const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
        const int64_t p0 = tile_offset(p, i);
        int64_t valid = a->n_paths - p0;
        if (valid > kTile) valid = kTile;
        const uint32_t elems = (uint32_t)((valid + 31) / 32 * 32);
        const int s = (int)(item % kStages);
        unsigned char* st = ring + (size_t)s * kStage;
        UT* const Ug = static_cast<UT*>(a->U);
        const uint32_t bu = elems * (uint32_t)sizeof(UT);
        const bool need_u = (p != 0);
        if (LEAN) {
            const uint32_t bl = elems * 4u;
            if (parts & 1) mbar_expect_tx(&full[s], bl + (need_u ? bu : 0u));
            if (parts & 2) {
                tma_load_1d_hint(st, a->L + p0, bl, &full[s], pol_keep);
                if (need_u) tma_load_1d_hint(st + kUOff, Ug + p0, bu, &full[s], pol_keep);
            }
        } else {
            const uint32_t bx = elems * (uint32_t)sizeof(XT);
            const bool dec = need_dec(p), mom = moments(p);
            const size_t col_bytes = (size_t)a->ld * sizeof(XT);
            const XT* xdec = reinterpret_cast<const XT*>(static_cast<const char*>(a->S) + (size_t)pass_t(p) * col_bytes);
            const XT* xreg = reinterpret_cast<const XT*>(static_cast<const char*>(a->S) + (size_t)(pass_t(p) - 1) * col_bytes);
            if (parts & 1) {
                mbar_expect_tx(&full[s], (dec ? bx : 0u) + (mom ? bx : 0u) + (need_u ? bu : 0u));
                if (dec) tma_load_1d_hint(st, xdec + p0, bx, &full[s], pol_stream);
                if (mom) tma_load_1d_hint(st + kTile * sizeof(XT), xreg + p0, bx, &full[s], pol_keep);
            }
            if ((parts & 2) && need_u) tma_load_1d_hint(st + kUOff, Ug + p0, bu, &full[s], pol_keep);
        }
    }
};

// One pass of one block: decide(t) + moments(t-1) over the block's tiles, then the block's partial row.  Out of line on
// purpose: inside this function only the streaming loop's state is live, so it gets the register allocation of the
// per-launch kernel of round 1 (no spills in the loop); the sweep-level state stays with the caller.
// `pre` = items of this pass whose column copies thread 0 issued at the end of the previous pass.
template <typename XT, typename UT, int D, bool LEAN>
__device__ __noinline__ void sweep_pass(const SweepArgs& a, unsigned char* ring, uint64_t* full, double* red, int my_tiles,
                                        int p, uint32_t item_base, int pre) {
    constexpr int NACC = 3 * D + 1;
    constexpr int kStage = SweepStage<XT, UT, LEAN>::value;
    constexpr int kUOff = SweepStage<XT, UT, LEAN>::kUOff;
    using U2 = typename Vec2<UT>::type;
    using V2 = typename Vec2<XT>::type;
    SweepRing<XT, UT, LEAN> rg{&a, ring, full, my_tiles};
    UT* const Ug = static_cast<UT*>(a.U);
    const uint64_t pol_keep = l2_policy_evict_last();

    const int t = a.n_steps - p;
    const int mode = (p == 0) ? kMaturity : (a.american ? kDecide : kObserve);
    const bool moments = rg.moments(p);
    const bool need_dec = LEAN ? true : (mode != kObserve);
    const bool need_u_in = (p != 0);
    const bool write_u = (mode != kObserve);

    if (threadIdx.x == 0) {
        // the state of the prefetched items (this block's own writes of the previous pass are ordered by the barriers
        // and proxy fences at the end of that pass); items beyond `pre` are issued whole inside the loop
        for (int i = 0; i < pre; ++i) rg.issue(p, i, item_base + (uint32_t)i, 2);
    }
    const SweepTab td = a.tab[t];
    const SweepTab tr = moments ? a.tab[t - 1] : td;
    double gam[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) gam[i] = (mode == kDecide) ? __ldcg(a.gamma + (size_t)t * kMaxK + i) : 0.0;

    StepArgs sa = {};                        // the launch-uniform view path_step() expects
    sa.t_dec = t;
    sa.mode = mode;
    sa.moments = moments ? 1 : 0;
    sa.is_put = a.is_put;
    sa.K = a.K;
    sa.disc_dec = td.disc;
    sa.mu_dec = td.mu; sa.isg_dec = td.isg;
    sa.mu_reg = tr.mu; sa.isg_reg = tr.isg;
    sa.first_hit = a.first_hit;
    sa.tau = a.tau;
    FastConsts fc;
    fc.sgn = a.is_put ? -1.0 : 1.0;
    fc.sgnK = a.is_put ? a.K : -a.K;
    fc.da = td.isg; fc.db = -td.mu * td.isg;
    fc.ra = tr.isg; fc.rb = -tr.mu * tr.isg;
    fc.disc = td.disc;
    const bool fast_ok = (mode == kDecide) && moments && !a.first_hit && !a.tau;

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    constexpr int NK = kTile / 2 / kStepThreads;
    for (int i = 0; i < my_tiles; ++i) {
        const uint32_t item = item_base + (uint32_t)i;
        // keep kStages - 1 items in flight
        const int want = i + kStages - 1;
        if (threadIdx.x == 0 && want < my_tiles && want >= pre) rg.issue(p, want, item_base + (uint32_t)want, 3);
        const int s = (int)(item % kStages);
        mbar_wait(&full[s], (item / kStages) & 1u);
        const int64_t p0 = rg.tile_offset(p, i);
        const unsigned char* st = ring + (size_t)s * kStage;

        if (LEAN) {
            // one quad of four adjacent paths per thread: prices of column t from L_t, increments of step t from the
            // counter, prices of column t-1 from L_{t-1} = L_t - q_t (written back as the new state)
            const int e0 = 4 * threadIdx.x;
            int64_t valid64 = a.n_paths - p0;
            const int valid = (int)(valid64 > kTile ? kTile : valid64);
            if (e0 < valid) {
                const int4 Lv = reinterpret_cast<const int4*>(st)[threadIdx.x];
                int Lt[4] = {Lv.x, Lv.y, Lv.z, Lv.w};
                double u[4] = {0.0, 0.0, 0.0, 0.0};
                if (need_u_in) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) u[j] = (double)reinterpret_cast<const UT*>(st + kUOff)[e0 + j];
                }
                int Lr[4] = {Lt[0], Lt[1], Lt[2], Lt[3]};
                if (moments) {
                    const uint64_t quad = (uint64_t)(a.quad0 + ((p0 + e0) >> 2));
                    int q[4];
                    if (a.rounds == 7) quad_increments<7>(a.gen, (uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)t, q);
                    else quad_increments<10>(a.gen, (uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)t, q);
#pragma unroll
                    for (int j = 0; j < 4; ++j) Lr[j] = Lt[j] - q[j];
                }
                float xdf[4], xrf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    xdf[j] = price_from_log(a.gen, Lt[j]);
                    xrf[j] = moments ? price_from_log(a.gen, Lr[j]) : xdf[j];
                }
                if (fast_ok && valid == kTile) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (fast_path_step<D>(fc, gam, (double)xdf[j], (double)xrf[j], u[j], true, acc)) Ug[p0 + e0 + j] = (UT)u[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int tj = 0, fj = 0;
                        const int64_t pj = p0 + e0 + j;
                        if (e0 + j < valid) {
                            if (a.first_hit) fj = __ldg(a.first_hit + pj);
                            if (a.tau && need_u_in) tj = a.tau[pj];
                            const bool ch = path_step<D>(sa, gam, (double)xdf[j], (double)xrf[j], u[j], tj, fj, acc);
                            if (write_u && ch) {
                                Ug[pj] = (UT)u[j];
                                if (a.tau) a.tau[pj] = tj;
                            }
                        }
                    }
                }
                if (moments) *reinterpret_cast<int4*>(a.L + p0 + e0) = make_int4(Lr[0], Lr[1], Lr[2], Lr[3]);
            }
        } else {
            const V2* sxd = reinterpret_cast<const V2*>(st) + threadIdx.x;
            const V2* sxr = reinterpret_cast<const V2*>(st + kTile * sizeof(XT)) + threadIdx.x;
            const U2* su = reinterpret_cast<const U2*>(st + kUOff) + threadIdx.x;
            if (fast_ok && p0 + kTile <= a.n_paths) {
                UT* const up = Ug + p0 + 2 * threadIdx.x;
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const V2 vd = sxd[k * kStepThreads], vr = sxr[k * kStepThreads];
                    const U2 uv = su[k * kStepThreads];
                    double2 u = make_double2((double)uv.x, (double)uv.y);
                    bool changed = fast_path_step<D>(fc, gam, (double)vd.x, (double)vr.x, u.x, true, acc);
                    changed |= fast_path_step<D>(fc, gam, (double)vd.y, (double)vr.y, u.y, true, acc);
                    // the state is written only where a path exercised (one vector store per pair)
                    if (changed) store_state2(up + 2 * k * kStepThreads, u, pol_keep);
                }
            } else {
                int64_t valid64 = a.n_paths - p0;
                const int valid = (int)(valid64 > kTile ? kTile : valid64);
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const int e0 = 2 * (threadIdx.x + k * kStepThreads);
                    if (e0 < valid) {
                        const bool two = (e0 + 1 < valid);
                        double xd0 = 0, xd1 = 0, xr0 = 0, xr1 = 0;
                        double2 u = make_double2(0.0, 0.0);
                        int2 f = make_int2(0, 0), tt = make_int2(0, 0);
                        if (need_dec) { const V2 v = sxd[k * kStepThreads]; xd0 = (double)v.x; xd1 = (double)v.y; }
                        if (moments) { const V2 v = sxr[k * kStepThreads]; xr0 = (double)v.x; xr1 = (double)v.y; }
                        if (need_u_in) { const U2 uv = su[k * kStepThreads]; u = make_double2((double)uv.x, (double)uv.y); }
                        const int64_t pp = p0 + e0;
                        if (a.first_hit) { f.x = __ldg(a.first_hit + pp); if (two) f.y = __ldg(a.first_hit + pp + 1); }
                        if (a.tau && need_u_in) { tt.x = a.tau[pp]; if (two) tt.y = a.tau[pp + 1]; }
                        bool changed = path_step<D>(sa, gam, xd0, xr0, u.x, tt.x, f.x, acc);
                        if (two) changed |= path_step<D>(sa, gam, xd1, xr1, u.y, tt.y, f.y, acc);
                        if (write_u && changed) {
                            if (two) {
                                store_state2(Ug + pp, u, pol_keep);
                                if (a.tau) *reinterpret_cast<int2*>(a.tau + pp) = tt;
                            } else {
                                Ug[pp] = (UT)u.x;
                                if (a.tau) a.tau[pp] = tt.x;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();                     // every warp is done with stage s before it is refilled
    }
    // this thread's state stores of the pass must be visible to the bulk-copy engine that re-reads them in the next pass
    // (generic -> async proxy); once per pass, ordered before thread 0's copies by the barriers that follow
    if (write_u || LEAN) fence_proxy_async_global();
    block_reduce_store<NACC, kStepThreads, kAccStride>(acc, red, a.partials + (int64_t)blockIdx.x * kAccStride);
    __syncthreads();
}

template <typename XT, typename UT, int D, bool LEAN>
__global__ void __launch_bounds__(kStepThreads, SweepMinBlocks<XT, D, LEAN>::value) lsm_sweep_kernel(const __grid_constant__ SweepArgs a) {
    constexpr int NACC = 3 * D + 1;
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ double red[(kStepThreads / 32) * NACC];
    __shared__ uint64_t full[kStages];
    __shared__ int s_alive, s_last;

    const int64_t n_tiles = (a.n_paths + kTile - 1) / kTile;
    const int my_tiles = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);     // static ownership
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        s_alive = 1;
    }
    __syncthreads();
    SweepRing<XT, UT, LEAN> rg{&a, ring, full, my_tiles};

    uint32_t item_base = 0;                 // ring items consumed before the current pass (uniform)
    int pre = min(kStages - 1, my_tiles);   // items of the current pass whose column copies are already issued
    if (threadIdx.x == 0)
        for (int i = 0; i < pre; ++i) rg.issue(0, i, (uint32_t)i, 1);

    for (int p = 0; p < a.n_passes; ++p) {
        // the polynomial of step t must be published (which also means: every block's row of the previous pass was read)
        if (p > 0) {
            if (threadIdx.x == 0 && !spin_until_at_least(a.sync + kSyncPublished, (uint32_t)p, a.sync)) s_alive = 0;
            __syncthreads();
            if (!s_alive) {
                // aborted: complete the copies already armed on this block's barriers before the block goes away
                if (threadIdx.x == 0) {
                    for (int i = 0; i < pre; ++i) rg.issue(p, i, item_base + (uint32_t)i, 2);
                    for (int i = 0; i < pre; ++i) {
                        const uint32_t item = item_base + (uint32_t)i;
                        mbar_wait(&full[item % kStages], (item / kStages) & 1u);
                    }
                }
                return;
            }
        }
        sweep_pass<XT, UT, D, LEAN>(a, ring, full, red, my_tiles, p, item_base, pre);
        item_base += (uint32_t)my_tiles;
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = (atomicAdd(a.sync + kSyncTickets + p, 1u) == gridDim.x - 1u) ? 1 : 0;
            // next pass: its first columns go out now, while the solve runs
            if (p + 1 < a.n_passes)
                for (int i = 0; i < pre; ++i) rg.issue(p + 1, i, item_base + (uint32_t)i, 1);
        }
        __syncthreads();
        if (s_last) {
            // every block's row is in: this block reduces them, exchanges with the peers and solves for step t-1
            __threadfence();
            sweep_solve_pass<D + 1>(a, p, a.n_steps - p);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
template <typename XT, typename UT, int D, bool LEAN>
static cudaError_t launch_sweep_t(int grid, const SweepArgs& a, cudaStream_t s) {
    constexpr int smem = kStages * SweepStage<XT, UT, LEAN>::value;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lsm_sweep_kernel<XT, UT, D, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(lsm_sweep_kernel<XT, UT, D, LEAN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    // cooperative: the launch fails (instead of deadlocking) if the grid could not be resident all at once
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kStepThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, lsm_sweep_kernel<XT, UT, D, LEAN>, a);
}

template <typename XT, typename UT, bool LEAN>
static cudaError_t launch_sweep_d(int degree, int grid, const SweepArgs& a, cudaStream_t s) {
    switch (degree) {
        case 0: return launch_sweep_t<XT, UT, 0, LEAN>(grid, a, s);
        case 1: return launch_sweep_t<XT, UT, 1, LEAN>(grid, a, s);
        case 2: return launch_sweep_t<XT, UT, 2, LEAN>(grid, a, s);
        case 3: return launch_sweep_t<XT, UT, 3, LEAN>(grid, a, s);
        case 4: return launch_sweep_t<XT, UT, 4, LEAN>(grid, a, s);
        case 5: return launch_sweep_t<XT, UT, 5, LEAN>(grid, a, s);
        case 6: return launch_sweep_t<XT, UT, 6, LEAN>(grid, a, s);
        case 7: return launch_sweep_t<XT, UT, 7, LEAN>(grid, a, s);
        case 8: return launch_sweep_t<XT, UT, 8, LEAN>(grid, a, s);
        case 9: return launch_sweep_t<XT, UT, 9, LEAN>(grid, a, s);
        case 10: return launch_sweep_t<XT, UT, 10, LEAN>(grid, a, s);
    }
    return cudaErrorInvalidValue;
}

template <typename XT, typename UT, int D, bool LEAN>
static int sweep_occupancy_t() {
    int nb = 0;
    constexpr int smem = kStages * SweepStage<XT, UT, LEAN>::value;
    cudaFuncSetAttribute(lsm_sweep_kernel<XT, UT, D, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(lsm_sweep_kernel<XT, UT, D, LEAN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lsm_sweep_kernel<XT, UT, D, LEAN>, kStepThreads, smem);
    return nb;
}

template <typename XT, typename UT, bool LEAN>
static int sweep_occupancy_d(int degree) {
    switch (degree) {
        case 0: return sweep_occupancy_t<XT, UT, 0, LEAN>();
        case 1: return sweep_occupancy_t<XT, UT, 1, LEAN>();
        case 2: return sweep_occupancy_t<XT, UT, 2, LEAN>();
        case 3: return sweep_occupancy_t<XT, UT, 3, LEAN>();
        case 4: return sweep_occupancy_t<XT, UT, 4, LEAN>();
        case 5: return sweep_occupancy_t<XT, UT, 5, LEAN>();
        case 6: return sweep_occupancy_t<XT, UT, 6, LEAN>();
        case 7: return sweep_occupancy_t<XT, UT, 7, LEAN>();
        case 8: return sweep_occupancy_t<XT, UT, 8, LEAN>();
        case 9: return sweep_occupancy_t<XT, UT, 9, LEAN>();
        case 10: return sweep_occupancy_t<XT, UT, 10, LEAN>();
    }
    return 1;
}

}  // namespace amc
