/* libamc -- C ABI of the B200-native Longstaff-Schwartz hot path.
 *
 * Drop-in boundary for /root/reference/american_monte_carlo.py:72-197 ("amc.py").  The reference has no
 * FFI of its own: its boundary is the Python module surface (`generate_asset_paths`,
 * `lsmc_option_pricing`, `intrinsic_value`, `regression_estimate`, ... imported by unit_test.py:3 and
 * american_monte_carlo_additional_plots.py:3).  Each entry point below names the reference function(s)
 * it replaces; `american_monte_carlo_b200/api.py` is the ctypes host side that keeps the reference's
 * Python signatures, and INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on failure; amc_last_error() gives the message of the
 *     last failure on the calling thread.  AMC_ERR_VALUE maps to Python ValueError, the rest to RuntimeError.
 *   - plain pointers and sizes only; `const double*` arguments are HOST pointers unless the name ends in
 *     `_dev`.  Opaque handles own device memory; free them with the matching *_free / *_destroy.
 *   - one amc_ctx per process and GPU (one process per GPU under torchrun); no internal locking.
 *   - path matrices live on the device TIMESTEP-MAJOR: S[t][p], t = 0..n, each column 128-byte aligned.
 *   - there is no CPU implementation behind any entry point: without a CUDA device amc_ctx_create fails.
 */
#ifndef AMC_H_
#define AMC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMC_OK 0
#define AMC_ERR_VALUE 1      /* bad argument (unknown basis, degree out of range, shape mismatch) */
#define AMC_ERR_CUDA 2       /* CUDA runtime failure */
#define AMC_ERR_NCCL 3       /* NCCL failure / NCCL not loadable */
#define AMC_ERR_STATE 4      /* handle used in the wrong state */

#define AMC_MAX_DEGREE 10
#define AMC_MAX_K (AMC_MAX_DEGREE + 1)

/* storage type of the path matrix */
#define AMC_F64 0
#define AMC_F32 1

/* basis ids: amc.py:99-101 has Power / Chebyshev / Legendre; Laguerre is an addition (BASELINE.json config 5) */
#define AMC_BASIS_POWER 0
#define AMC_BASIS_CHEBYSHEV 1
#define AMC_BASIS_LEGENDRE 2
#define AMC_BASIS_LAGUERRE 3

typedef struct amc_ctx amc_ctx;
typedef struct amc_paths amc_paths;

const char* amc_last_error(void);
int amc_version(void);

/* ---- context ------------------------------------------------------------------------------------------ */
/* `stream` is a cudaStream_t to enqueue all work on (e.g. torch.cuda.current_stream().cuda_stream), or NULL to
 * let the library create its own non-blocking stream. */
int amc_ctx_create(int device, void* stream, amc_ctx** out);
int amc_ctx_destroy(amc_ctx* ctx);
int amc_ctx_sync(amc_ctx* ctx);
int amc_ctx_device_info(amc_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem);

/* ---- multi-GPU: one process per GPU, paths sharded, one all-reduce of the moment sums per step -------- */
/* 128-byte ncclUniqueId; rank 0 creates it, the host side broadcasts it (torch.distributed) to all ranks. */
int amc_comm_unique_id(char id[128]);
int amc_comm_init(amc_ctx* ctx, int world_size, int rank, const char id[128]);
int amc_comm_info(amc_ctx* ctx, int* world_size, int* rank);
/* how the per-step all-reduce of the moment sums travels: 0 = single GPU (none), 1 = ncclAllReduce between two solve
 * launches, 2 = fused into the solve kernel over NVLink peer memory (CUDA-IPC-mapped mailboxes; the default when
 * every rank can map every peer; AMC_ALLREDUCE=nccl|p2p forces one) */
int amc_comm_transport(amc_ctx* ctx, int* transport);
/* sum-all-reduce `n` doubles in place on a host buffer (tiny; used for column statistics and tests) */
int amc_comm_allreduce_host(amc_ctx* ctx, double* buf, int n);

/* ---- path simulation: replaces generate_asset_paths, amc.py:72-81 -------------------------------------- */
/* K1: counter-based Philox4x32-10 + Box-Muller, log-space cumulative sum, timestep-major vector stores.
 * `n_paths_local` paths whose GLOBAL ids start at `path_offset` (so the union over ranks does not depend on
 * the number of ranks); `n_paths_global` is the total over all ranks.
 * AMC_F32: one Philox call per quad of adjacent paths and step, log2-prices summed exactly in int32 fixed point
 * (csrc/gbm_quad.cuh); AMC_F64: 53-bit uniforms, double log-price.  AMC_PHILOX_ROUNDS=7 selects Philox4x32-7 for the
 * float generator (explicit throughput option; 10 rounds is the default). */
int amc_paths_generate(amc_ctx* ctx, double S0, double r, double sigma, double T, int n_time_steps,
                       int64_t n_paths_local, int64_t path_offset, int64_t n_paths_global, int dtype,
                       uint64_t seed, amc_paths** out);
/* Path-free ("lean") variant of amc_paths_generate for float paths: NO path matrix is stored --
 * only each path's terminal fixed-point log2-price (4 bytes per path instead of 4 (n+1)).  amc_lsm_price regenerates
 * column t-1 from column t inside the backward sweep (L_{t-1} = L_t - q_t, q_t recomputed from the Philox counter: the
 * exact reverse of the forward integer sum), so prices and exercise decisions equal those of the stored set with the
 * same seed.  `path_offset` must be a multiple of 4.  Read-back entry points (amc_paths_column / _rows / ...) work on a
 * lean set by walking the counters forward again (diagnostics; O(t) per element). */
int amc_paths_generate_lean(amc_ctx* ctx, double S0, double r, double sigma, double T, int n_time_steps,
                            int64_t n_paths_local, int64_t path_offset, int64_t n_paths_global, uint64_t seed,
                            amc_paths** out);
/* K1z: same arithmetic from caller-supplied standard normals Z[p][j] (row-major [n_paths_local][n_time_steps],
 * exactly what amc.py:74 draws) -- the A/B mode: identical inputs to the reference. */
int amc_paths_from_normals(amc_ctx* ctx, const double* Z, double S0, double r, double sigma, double T,
                           int n_time_steps, int64_t n_paths_local, int64_t n_paths_global, int dtype,
                           amc_paths** out);
/* same, Z already on the device (row-major, f64) */
int amc_paths_from_normals_dev(amc_ctx* ctx, const double* Z_dev, double S0, double r, double sigma, double T,
                               int n_time_steps, int64_t n_paths_local, int64_t n_paths_global, int dtype,
                               amc_paths** out);
/* adopt a path matrix computed elsewhere: S[p][t] row-major [n_paths_local][n_time_steps+1] (the layout
 * amc.py:78-81 returns); transposed on the device, per-column statistics measured there. */
int amc_paths_from_host(amc_ctx* ctx, const double* S, int n_time_steps, int64_t n_paths_local,
                        int64_t n_paths_global, int dtype, amc_paths** out);
int amc_paths_free(amc_paths* paths);
int amc_paths_info(const amc_paths* paths, int64_t* n_paths_local, int64_t* n_paths_global, int* n_time_steps,
                   int* dtype, int64_t* bytes_on_device);
/* read back (for __array__, plotting, tests): one column S[:, t] or rows [p0, p1) as [p1-p0][n+1] row-major */
int amc_paths_column(const amc_paths* paths, int t, double* out);
int amc_paths_rows(const amc_paths* paths, int64_t p0, int64_t p1, double* out);
/* per-column affine maps (mu_t, sigma_t) used to standardise S[:, t] inside the kernels */
int amc_paths_column_maps(const amc_paths* paths, double* mu, double* sigma);

/* ---- LSM backward induction: replaces lsmc_option_pricing / perform_backward_iteration /
 *      estimate_continuation_values / regression_estimate / apply_exercise / precompute_barrier_hit_matrix,
 *      amc.py:90-197 ------------------------------------------------------------------------------------- */
typedef struct amc_lsm_spec {
    double K;                /* strike */
    double r;                /* risk-free rate */
    double dt;               /* step length passed by the caller (amc.py:180) */
    double barrier;          /* down-and-in level; NaN = no barrier (amc.py:172 `is not None`) */
    double scaling_factor;   /* regression_estimate(scaling_factor=2) */
    int is_put;              /* option_type == "Put" (anything else is a call, amc.py:86) */
    int is_american;         /* exercise_type == "American" (anything else never exercises early, amc.py:154) */
    int basis;               /* AMC_BASIS_* */
    int degree;              /* 0..AMC_MAX_DEGREE */
    int scaling;             /* regression_estimate(scaling=False) */
    int want_regression;     /* run the per-step regressions even when they cannot change the price
                                (European exercise): needed for continuation values, amc.py:151,164 */
    int want_exercise_steps; /* keep a per-path exercise-step array on the device (tests / diagnostics) */
    int want_svd;            /* always run the k x k SVD and report singular values (otherwise it is skipped on
                                steps whose full rank is certified cheaply; prices do not depend on this) */
    int state_f32;           /* keep the per-path state (cashflow discounted to time 0) in float instead of double:
                                FP32-path mode only (AMC_F32 path sets); all sums, the solve and the exercise test stay
                                in double.  Halves the state traffic (2 b_S + 8 instead of 2 b_S + 16 bytes per
                                path-step); price effect ~1e-8 relative, inside the 1e-5 FP32 tolerance */
} amc_lsm_spec;

/* per-step diagnostics, all indexed by t = 0..n_time_steps (entry n is unused: no regression at maturity) */
typedef struct amc_lsm_steps {
    double* gamma;   /* [(n+1)][AMC_MAX_K] continuation polynomial in z = (x - mu_t)/sigma_t (internal basis) */
    double* beta;    /* [(n+1)][AMC_MAX_K] numpy-lstsq-equivalent coefficients in the user's basis */
    double* sv;      /* [(n+1)][AMC_MAX_K] singular values of the design matrix, descending (rows where the
                        SVD ran: always with spec.want_svd, else only on steps that were not certified) */
    double* mean_x;  /* [(n+1)] np.mean(paths[:, t]) */
    double* std_x;   /* [(n+1)] np.std(paths[:, t]) */
    int* rank;       /* [(n+1)] numpy's rank */
    double* pivot_loss; /* [(n+1)] conditioning report of the moment-based solve: max over the Cholesky pivots of the
                        internal Gram of G_jj / pivot_j (>= 1, ~cond(G); 0 where no regression ran).  Up to ~1e11 the
                        rank decision and the fit reproduce numpy.linalg.lstsq (tests/, scripts/explore_rank.py); from
                        ~1e12 on -- degree 10 on very heavy-tailed columns, sigma*sqrt(T) ~ 1 -- they may deviate:
                        lower the degree (the Python host warns) */
} amc_lsm_steps;     /* any pointer may be NULL */

typedef struct amc_lsm_timing {
    float total_ms;        /* whole backward sweep on the device (CUDA events on the context stream) */
    float step_kernel_ms;  /* sum over the fused decide+moments launches (only when profile != 0) */
    float solve_kernel_ms; /* sum over the solve launches incl. all-reduce (only when profile != 0) */
    int step_launches;
    int solve_launches;
    int other_launches;
    int sweep_kind;        /* which kernels ran the sweep: 0 = per-step launch chain, 1 = persistent cooperative kernel
                              (path-free sets, AMC_PERSISTENT=1), 2 = one-cluster kernel (small stored sets) */
} amc_lsm_timing;

/* Price one contract on a device-resident path set.  `price` is the GLOBAL mean over all ranks' paths.
 * exercise_step_out (optional, host, [n_paths_local] int32) needs spec->want_exercise_steps;
 * cashflow0_out (optional, host, [n_paths_local]) = each path's cashflow discounted to time 0. */
int amc_lsm_price(amc_ctx* ctx, const amc_paths* paths, const amc_lsm_spec* spec, double* price,
                  amc_lsm_steps* steps, int32_t* exercise_step_out, double* cashflow0_out,
                  amc_lsm_timing* timing, int profile);

/* amc_lsm_price with the knock-in information supplied by the caller instead of a barrier level: first_hit[p] = first
 * step at which path p counts as knocked in (n_time_steps+1 = never) -- the running-OR matrix that
 * perform_backward_iteration (amc.py:139-167) receives, as one index per path.  spec->barrier is ignored. */
int amc_lsm_price_with_hits(amc_ctx* ctx, const amc_paths* paths, const amc_lsm_spec* spec, const int32_t* first_hit,
                            double* price, amc_lsm_steps* steps, int32_t* exercise_step_out, double* cashflow0_out,
                            amc_lsm_timing* timing, int profile);
/* out[p] = S[p, steps[p]] (host arrays, [n_paths_local]): e.g. each path's price at its exercise step, from which the
 * undiscounted cashflows of amc.py:93,148 follow exactly. */
int amc_paths_gather_steps(const amc_paths* paths, const int32_t* steps, double* out);

/* Price `n_contracts` contracts on ONE device-resident path set in the same launches (BASELINE.json config 4: the
 * strike axis of a strike x vol x maturity grid shares its paths; the reference re-simulates and re-prices per
 * contract in a Python loop, american_monte_carlo_additional_plots.py:100-107).  The contracts may differ in K,
 * is_put and is_american; r, dt, barrier, basis, degree, scaling and scaling_factor must equal those of specs[0].
 * Each contract's result is the one amc_lsm_price gives for it alone (same arithmetic; the partial sums are grouped
 * differently, so agreement is to rounding, not bitwise).  prices: [n_contracts]; gamma_out (optional):
 * [n_contracts][n_time_steps+1][AMC_MAX_K].  Multi-GPU use shards the CONTRACTS (and their path sets) over ranks:
 * the path set passed here must be complete (n_paths_local == n_paths_global); no collective is involved. */
#define AMC_MAX_BATCH 256
int amc_lsm_price_batch(amc_ctx* ctx, const amc_paths* paths, const amc_lsm_spec* specs, int n_contracts,
                        double* prices, double* gamma_out, amc_lsm_timing* timing, int profile);

/* Continuation value max(fit, 0) of every local path at step t from a stored gamma (lazy replacement for the
 * per-step copies of amc.py:164).  out: host [n_paths_local]. */
int amc_continuation(amc_ctx* ctx, const amc_paths* paths, int t, const double* gamma, int degree, double* out);

/* ---- exposures: replaces compute_ccr_exposures, amc.py:400-414 ------------------------------------------------ */
/* Per step t = 0..n: the q_lo / q_hi quantiles (numpy's default `linear` rule, q in [0, 1]; the reference uses 0.05 and
 * 0.95) and the mean of the finite continuation values max(fit_t, 0) of ALL paths (global over ranks for a sharded path
 * set), computed from the stored continuation polynomials `gamma` ([n+1][AMC_MAX_K], as returned in amc_lsm_steps.gamma)
 * without materialising the [n_paths] vectors; step n is all zeros (amc.py:145).  Outputs: host arrays [n+1]. */
int amc_ccr_exposures(amc_ctx* ctx, const amc_paths* paths, const double* gamma, int degree, double q_lo, double q_hi,
                      double* pfe_lo, double* pfe_hi, double* epe);
/* The same three numbers (q_lo quantile, q_hi quantile, mean of the finite values) for an arbitrary host array:
 * compute_ccr_exposures on values that did not come from this library (amc.py:478 feeds it QuantLib values).
 * out3 = NaN, NaN, NaN when no value is finite (amc.py:405-408). */
int amc_percentiles(amc_ctx* ctx, const double* values, int64_t n, double q_lo, double q_hi, double out3[3]);

/* ---- small array ops kept for API parity (host arrays in, host arrays out, computed on the device) ---- */
/* intrinsic_value, amc.py:85-86 */
int amc_intrinsic_value(amc_ctx* ctx, const double* S, int64_t n, double K, int is_put, double* out);
/* regression_estimate, amc.py:110-122: fitted values of the (possibly rank-truncated) least-squares fit */
int amc_regression_fit(amc_ctx* ctx, const double* X, const double* Y, int64_t n, int basis, int degree,
                       int scaling, double scaling_factor, double* fitted, double* beta, int* rank);
/* estimate_continuation_values, amc.py:126-135, on host arrays: Y = cashflows * exp(-r dt (exercise_times - t)) is formed
 * on the device, regressed on X (= paths[:, t]) and the fitted values are clamped at zero (amc.py:132) */
int amc_estimate_continuation(amc_ctx* ctx, const double* X, const double* cashflows, const int64_t* exercise_times,
                              int64_t n, int64_t t, double r, double dt, int basis, int degree, int scaling,
                              double scaling_factor, double* out);
/* apply_exercise, amc.py:90-94: for i < m with exercise_value[i] > continuation[i] (strict) set
 * cashflows[indices[i]] = exercise_value[i] and exercise_times[indices[i]] = t; both arrays ([n_total]) are updated in place */
int amc_apply_exercise(amc_ctx* ctx, double* cashflows, int64_t* exercise_times, int64_t n_total,
                       const double* exercise_value, const double* continuation, const int64_t* indices, int64_t m,
                       int64_t t);
/* get_basis_polynomials, amc.py:98-106: out is [n][degree+1] row-major */
int amc_basis_matrix(amc_ctx* ctx, const double* X, int64_t n, int basis, int degree, double* out);
/* precompute_barrier_hit_matrix, amc.py:171-176: out is [n_paths][n_time_steps+1] row-major bytes (0/1) */
int amc_barrier_hit_matrix(amc_ctx* ctx, const amc_paths* paths, double barrier, uint8_t* out);

/* ---- generator self-tests (diagnostics; no reference counterpart: amc.py:74 uses NumPy's host stream) -------------- */
/* The DEVICE build of Philox4x32-`rounds` (10 or 7) on `n` caller-chosen counters: out[i] = philox(counters[i], key).
 * The Random123 known-answer vectors must come out bit-exact (tests/test_gpu_generator.py). */
int amc_selftest_philox(amc_ctx* ctx, int rounds, const uint32_t* counters, const uint32_t key[2], int n, uint32_t* out);
/* Histogram of the float generator's standard normals, formed exactly as the path kernel forms them (Philox quad
 * counters (quad, step), Box-Muller on the MUFU approximations): 4 * n_quads * n_steps samples binned on [lo, hi) into
 * n_bins (<= 4096) equal bins; hist[0] / hist[n_bins+1] = under / overflow.  stats = count, sum z, sum z^2, sum z^3,
 * sum z^4, max |z|. */
int amc_selftest_normals(amc_ctx* ctx, int rounds, uint64_t seed, int64_t n_quads, int n_steps, int n_bins, double lo,
                         double hi, uint64_t* hist, double stats[6]);

#ifdef __cplusplus
}
#endif
#endif /* AMC_H_ */
