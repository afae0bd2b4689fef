// Host-callable launchers of the device kernels (implemented in pathgen.cu / lsm_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gbm_quad.cuh"
#include "lsm_solve.h"

namespace amc {

constexpr int kAccStride = 32;     // doubles per partial-sum row: 2d moment sums + (d+1) cross sums <= 31
constexpr int kStepThreads = 256;
constexpr int kPadElems = 32;      // columns are padded to a multiple of 32 elements (128 B for f32)

inline int64_t padded_len(int64_t n) { return (n + kPadElems - 1) / kPadElems * kPadElems; }

enum StepMode {
    kMaturity = 0,   // t = n: cashflow = payoff where knocked in (amc.py:147-149)
    kDecide = 1,     // t < n, American: exercise where payoff > max(fit, 0) (amc.py:154-162, 90-94)
    kObserve = 2     // t < n, no early exercise: the state is only read (regression target for t-1)
};

// Contract batches (BASELINE.json config 4): several contracts priced on ONE path set in the same launches.
// blockIdx.y selects the contract; its state vector, continuation polynomial and partial-sum rows are found through
// the strides below, strike / payoff side / exercise style through this table (device memory).
struct BatchContract {
    double K;
    int is_put;
    int is_american;
};

struct StepArgs {
    const void* x_dec;         // column t_dec   (null when nothing is decided: kObserve)
    const void* x_reg;         // column t_dec-1 (null when moments == 0)
    void* U;                   // per-path cashflow discounted to time 0 (double, or float with state_f32)
    int32_t* tau;              // optional exercise step per path
    const int32_t* first_hit;  // optional first knock-in step per path (barrier)
    const double* coef;        // gamma of step t_dec (device, kMaxK doubles); unused at maturity
    double* partials;          // [grid][kAccStride]
    int64_t n_paths;           // local
    int t_dec;
    int mode;
    int moments;               // 1: moment sums of column t_dec-1; 0: only sum(U) (last launch -> price)
    int is_put;
    int reverse;               // walk the tiles from the high end: consecutive launches alternate direction so the
                               // tail of what launch t touched last (S_{t-1}, U) is still in L2 when launch t-1 starts
    double K, disc_dec;        // strike, exp(-r dt t_dec)
    double mu_dec, isg_dec;    // affine map of column t_dec:   z = (x - mu) * isg
    double mu_reg, isg_reg;    // affine map of column t_dec-1
    // batch (null / 0 for a single contract): contract c = blockIdx.y uses U + c*u_stride, coef + c*coef_stride,
    // partial rows [c*gridDim.x, (c+1)*gridDim.x)
    const BatchContract* batch;
    int64_t u_stride;
    int64_t coef_stride;
};

// One-shot all-reduce of the moment sums over NVLink peer memory, fused into the solve kernel (multi-GPU).
// Every rank owns a mailbox of kPeerRing x world x kAccStride 16-byte cells {lo32, seq, hi32, seq}; the solve block of
// rank r stores its reduced sums straight into slot [seq % kPeerRing][r] of EVERY rank's mailbox (peer stores through
// NVSwitch), then spins on its own mailbox until all `world` rows carry the current sequence number and adds them in
// rank order -- the same bits on every rank, no NCCL kernel, no extra launch, no fence (the flag travels with the data,
// as in NCCL's LL protocol).
constexpr int kPeerMax = 8;
constexpr int kPeerRing = 4;
struct PeerArgs {
    uint4* mailbox[kPeerMax] = {};  // mailbox[q] = rank q's mailbox mapped into this process (cudaIpc); [rank] = own
    int world = 0, rank = 0;        // world <= 1: no exchange
    uint32_t* seq_ctr = nullptr;    // launch chain: device counter from which every solve launch takes the next sequence
                                    // number (never 0), so the launch arguments stay constant from sweep to sweep
                                    // (CUDA-graph replay); the host re-seeds it at the start of every sharded sweep with
                                    // its own count of exchanges, which keeps the ranks in lockstep whatever happened to
                                    // an earlier sweep.  The persistent kernel gets its numbers as SweepArgs::seq_base.
    int* err = nullptr;             // device flag: set to 1 if a peer did not show up in time
};

struct SolveArgs {
    const double* partials;    // [n_rows][kAccStride] (null: skip the reduction, sums already hold totals)
    int n_rows;
    double* sums;              // [kAccStride] reduced (and, multi-GPU, all-reduced) sums
    int do_reduce, do_solve, final_price;
    SolveSpec spec;
    double y_scale;            // exp(r dt t): brings the time-0 cashflows to time t
    double mu_ref, sigma_ref;  // affine map of the regressed column
    // outputs (device)
    double* gamma;             // [kMaxK]
    double* beta;              // [kMaxK]
    double* sv;                // [kMaxK]
    double* mean_std;          // [3]: mean, std of the column, pivot loss of the Gram factorisation
    int* rank;                 // [1]
    double* price;             // [1] (final_price)
    // batch: block c = blockIdx.x handles contract c: partials + c*n_rows rows, sums + c*kAccStride,
    // gamma + c*gamma_stride, price + c (beta / sv / mean_std / rank are not reported for batches)
    int n_batch;
    int64_t gamma_stride;
    PeerArgs peer;
};

// ---- persistent one-launch sweep (lsm_sweep.cuh) ---------------------------------------------------------------
struct SweepTab {             // per column t
    double disc;              // exp(-r dt t)
    double mu, isg;           // affine map  z = (x - mu) * isg
    double pad;
};

// sync words (uint32), each group on its own 128-byte line
constexpr int kSyncPublished = 0;      // number of solves published
constexpr int kSyncAbort = 32;         // nonzero: somebody timed out, everybody leaves
constexpr int kSyncTickets = 64;       // [n_passes]: blocks done with pass p
constexpr unsigned long long kSpinLimitNs = 8000000000ull;

struct SolverTab {            // per regressed column t
    double y_scale;           // exp(r dt t): brings the time-0 cashflows to time t
    double mu, sigma;         // affine map of the column
    double pad;
};

struct SweepArgs {
    const void* S;            // path matrix, timestep-major (column t at S + t * ld); null in LEAN mode
    int64_t ld;
    int32_t* L;               // LEAN: log2-price state, in: L_n, rewritten every pass
    void* U;
    int32_t* tau;
    const int32_t* first_hit;
    const SweepTab* tab;      // [n+1]
    const double* gamma;      // [n+1][kMaxK], row t written by the solver before pass n-t
    double* partials;         // [gridDim.x][kAccStride]
    uint32_t* sync;
    int64_t n_paths;
    int n_steps;
    int n_passes;             // n+1 with regressions, 1 without (price = discounted mean payoff)
    int american;
    int is_put;
    int reverse;              // alternate the tile direction from pass to pass (L2: the tail of pass p is the head of p+1)
    double K;
    QuadGen gen;              // LEAN
    int64_t quad0;            // LEAN: global quad id of local path 0 (path_offset / 4)
    int rounds;               // LEAN: Philox rounds the path set was generated with (10 or 7)
    // the solve of every pass is done by whichever block finishes the pass last (lsm_solve_block.cuh)
    SolveArgs solve;          // spec, partials, n_rows, sums, peer; gamma / beta / sv / mean_std / rank / price = TABLE bases
    const SolverTab* solve_tab;   // [n+1]
    uint32_t seq_base;        // multi-GPU: pass p exchanges under sequence number seq_base + p + 1
};


// grid of the cooperative launch: SM count x resident blocks per SM (every block resident: they wait on each other)
int sweep_grid_size(int dtype, int state_f32, int degree, int lean, int sm_count);
cudaError_t launch_sweep(int dtype, int state_f32, int degree, int lean, int grid, const SweepArgs& a, cudaStream_t s);


// one-cluster sweep for small path sets (lsm_cluster.cuh): the largest local path count it takes for this storage /
// state / degree on the current device (0: unavailable), and its launch (grid = one cluster)
int64_t cluster_sweep_capacity(int dtype, int state_f32, int degree);
cudaError_t launch_cluster_sweep(int dtype, int state_f32, int degree, const SweepArgs& a, cudaStream_t s);

int step_grid_size(int dtype, int state_f32, int degree, int sm_count);
// pdl: launch as a programmatic dependent of the previous kernel in the stream (see common.cuh)
cudaError_t launch_step(int dtype, int state_f32, int degree, int grid, const StepArgs& a, cudaStream_t s,
                        bool pdl = false, int n_batch = 1);
cudaError_t launch_solve(const SolveArgs& a, cudaStream_t s, bool pdl = false);
cudaError_t launch_continuation(int dtype, const void* x, int64_t n, const double* gamma_dev, int degree, double mu,
                                double isg, int clamp, double* out_dev, cudaStream_t s);
cudaError_t launch_discount(const double* cf_dev, const int64_t* tau_dev, int64_t n, int64_t t, double r, double dt,
                            double* y_dev, cudaStream_t s);
cudaError_t launch_apply_exercise(double* cf_dev, int64_t* tau_dev, const double* ev_dev, const double* cont_dev,
                                  const int64_t* idx_dev, int64_t m, int64_t t, cudaStream_t s);
cudaError_t launch_intrinsic(const double* S_dev, int64_t n, double K, int is_put, double* out_dev, cudaStream_t s);
cudaError_t launch_basis_matrix(const double* X_dev, int64_t n, int basis, int degree, double* out_dev,
                                cudaStream_t s);

// ccr.cu -- exposures (amc.py:400-414): radix select of the percentiles' order statistics + mean of the finite values
constexpr int kSelTargets = 4;      // floor / floor+1 order statistics of two percentiles
constexpr int kSelBins = 2048;      // 11 bits per pass
constexpr int kSelPasses = 6;       // 11 * 5 + 9 = 64 key bits
struct CcrSource {                  // where the values of one step come from
    const void* x;                  // path column (continuation value = clamp(poly((x - mu) * isg)))
    const double* vals;             // or: an explicit array (device)
    int x_f32, degree, clamp, zero; // zero: all values are 0 (maturity, amc.py:145)
    double mu, isg;
    double gam[kMaxK];
};
struct SelState {
    unsigned long long prefix[kSelTargets];   // key bits decided so far (right-aligned)
    unsigned long long rank[kSelTargets];     // rank among the keys that share the prefix
    unsigned long long count;                 // finite values
    double sum;                               // of the finite values
    double frac[2];                           // numpy's gamma of the two percentiles
    int bits_done;
};
cudaError_t launch_ccr_hist(const CcrSource& s, int64_t n, const SelState* st, int pass, unsigned long long* hist,
                            double* sum_partials, int grid, cudaStream_t stream);
cudaError_t launch_ccr_scan(SelState* st, unsigned long long* hist, int pass, const double* sum_partials, int n_partials,
                            double q_lo, double q_hi, double* out3_dev, cudaStream_t stream);

// pathgen.cu
struct GbmParams {
    double S0, drift, vol;     // per-step log drift (r - sigma^2/2) dt and vol sigma sqrt(dt)
};
QuadGen make_quad_gen(const GbmParams& g, int n_steps, uint64_t seed);
int philox_rounds();      // 10, or 7 with AMC_PHILOX_ROUNDS=7 (float generator only; philox.cuh)
cudaError_t launch_generate_philox(int dtype, void* S, int32_t* L_out, int64_t ld, int n_steps, int64_t n_local,
                                   int64_t path_offset, GbmParams g, uint64_t seed, int sm_count, cudaStream_t s);
// lean (path-free) sets: mode 0 = column t_stop -> out[n_local], 1 = rows [p_lo, p_hi) -> out[(p_hi-p_lo)][n+1],
// 2 = first knock-in step -> out_i[n_local], 3 = out[p] = S[steps[p]][p]; all by walking forward from the counters
cudaError_t launch_lean_walk(int mode, int rounds, const QuadGen& g, int64_t quad0, int64_t n_local, int n_steps, int t_stop,
                             int64_t p_lo, int64_t p_hi, double barrier, const int32_t* steps_dev, double* out_dev,
                             int32_t* out_i_dev, int sm_count, cudaStream_t s);
cudaError_t launch_philox_kat(int rounds, const uint32_t* ctr_dev, uint32_t k0, uint32_t k1, int n, uint32_t* out_dev,
                              cudaStream_t s);
cudaError_t launch_normals_hist(int rounds, uint64_t seed, int64_t n_quads, int n_steps, int n_bins, double lo, double hi,
                                unsigned long long* hist_dev, double* stats_dev, int grid, cudaStream_t s);
cudaError_t launch_from_normals(int dtype, const double* Z_dev, void* S, int64_t ld, int n_steps, int64_t n_local,
                                GbmParams g, cudaStream_t s);
cudaError_t launch_transpose_in(int dtype, const double* S_rowmajor_dev, void* S, int64_t ld, int n_cols,
                                int64_t n_local, cudaStream_t s);
cudaError_t launch_gather_rows(int dtype, const void* S, int64_t ld, int n_cols, int64_t p0, int64_t p1,
                               double* out_dev, cudaStream_t s);
cudaError_t launch_gather_steps(int dtype, const void* S, int64_t ld, int n_cols, int64_t n, const int32_t* steps_dev,
                                double* out_dev, cudaStream_t s);
cudaError_t launch_column_to_f64(int dtype, const void* col, int64_t n, double* out_dev, cudaStream_t s);
// shifted one-pass column statistics: partial[(col * n_chunks + chunk) * 2 + {0,1}] = sum(x - c), sum((x - c)^2)
// with c = first element of the column
cudaError_t launch_column_stats(int dtype, const void* S, int64_t ld, int n_cols, int64_t n_local, int n_chunks,
                                double* partial_dev, double* shift_dev, cudaStream_t s);
cudaError_t launch_first_hit(int dtype, const void* S, int64_t ld, int n_cols, int64_t n_local, double barrier,
                             int32_t* first_hit_dev, cudaStream_t s);
cudaError_t launch_hit_matrix(const int32_t* first_hit_dev, int n_cols, int64_t n_local, uint8_t* out_dev,
                              cudaStream_t s);

}  // namespace amc
