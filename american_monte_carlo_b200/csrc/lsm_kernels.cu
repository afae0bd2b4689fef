// Solve kernel, small array kernels and the launch dispatch of the LSM sweep (sm_100a).
// The fused decide+moments step kernels live in lsm_step.cuh (compiled per storage type in lsm_step_f32/f64.cu).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "lsm_solve_block.cuh"
#include "lsm_sweep.cuh"

namespace amc {

cudaError_t launch_step_f32(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch);
cudaError_t launch_step_f64(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch);
cudaError_t launch_step_f32s(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch);
int step_occupancy_f32(int degree);
int step_occupancy_f64(int degree);
int step_occupancy_f32s(int degree);
cudaError_t launch_sweep_f32(int degree, int grid, const SweepArgs& a, cudaStream_t s);
cudaError_t launch_sweep_f64(int degree, int grid, const SweepArgs& a, cudaStream_t s);
cudaError_t launch_sweep_f32s(int degree, int grid, const SweepArgs& a, cudaStream_t s);
cudaError_t launch_sweep_lean(int state_f32, int degree, int grid, const SweepArgs& a, cudaStream_t s);
int sweep_occupancy_f32(int degree);
int sweep_occupancy_f64(int degree);
int sweep_occupancy_f32s(int degree);
int sweep_occupancy_lean(int state_f32, int degree);
cudaError_t launch_cluster_f32(int degree, const SweepArgs& a, cudaStream_t s);
cudaError_t launch_cluster_f64(int degree, const SweepArgs& a, cudaStream_t s);
cudaError_t launch_cluster_f32s(int degree, const SweepArgs& a, cudaStream_t s);
int64_t cluster_capacity_f32(int degree);
int64_t cluster_capacity_f64(int degree);
int64_t cluster_capacity_f32s(int degree);

// ---------------------------------------------------------------------------------------------------------
// Solve kernel: <<<1, 128>>>.
//   phase 1 (do_reduce): sums[a] = sum over rows of partials[row][a], fixed order (lane-strided, then a
//                        shuffle tree) -> bitwise reproducible, no floating-point atomics anywhere.
//   phase 2 (do_solve):  thread 0 runs the k x k solve and stores gamma / diagnostics.
//   final_price:         price = sum(U) / P.
// Multi-GPU: phase 1, then an NCCL all-reduce of `sums`, then phase 2 as a second launch.
template <int K>
__global__ void __launch_bounds__(kSolveThreads, 1) lsm_solve_kernel(const SolveArgs a_in) {
    SolveArgs a = a_in;
    if (a_in.n_batch > 1) {                       // contract batches: block c owns contract c
        const int64_t c = blockIdx.x;
        a.partials = a_in.partials + c * a_in.n_rows * kAccStride;
        a.sums = a_in.sums + c * kAccStride;
        a.gamma = a_in.gamma + c * a_in.gamma_stride;
        a.price = a_in.price + c;
        a.beta = nullptr; a.sv = nullptr; a.mean_std = nullptr; a.rank = nullptr;
    }
    pdl_launch_dependents();
    pdl_wait();
    // multi-GPU: every exchange takes the next sequence number from a device counter (never 0), so the launch arguments of
    // a sweep stay constant from call to call (CUDA-graph replay); the host re-seeds the counter at the start of every
    // sharded sweep (api.cu), which keeps the ranks in lockstep whatever happened to an earlier sweep
    uint32_t seq = 0u;
    if (a.peer.world > 1) {
        __shared__ uint32_t s_seq;
        if (threadIdx.x == 0) {
            uint32_t v = atomicAdd(a.peer.seq_ctr, 1u) + 1u;
            if (v == 0u) v = atomicAdd(a.peer.seq_ctr, 1u) + 1u;
            s_seq = v;
        }
        __syncthreads();
        seq = s_seq;
    }
    solve_block<K, false>(a, a_in, seq, nullptr);
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

// Regression target of estimate_continuation_values (amc.py:128): Y = cashflows * exp(-r dt (tau - t)).
__global__ void discount_kernel(const double* __restrict__ cf, const int64_t* __restrict__ tau, int64_t n, int64_t t,
                                double r, double dt, double* __restrict__ y) {
    const double mrdt = -r * dt;                       // the reference's expression order: (-r * dt) * (tau - t)
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
        y[p] = cf[p] * exp(mrdt * (double)(tau[p] - t));
}

// apply_exercise (amc.py:90-94): where exercise_value > continuation (strict), scatter value and step.
__global__ void apply_exercise_kernel(double* __restrict__ cf, int64_t* __restrict__ tau, const double* __restrict__ ev,
                                      const double* __restrict__ cont, const int64_t* __restrict__ idx, int64_t m,
                                      int64_t t) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        if (ev[i] > cont[i]) {
            cf[idx[i]] = ev[i];
            tau[idx[i]] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// dtype: storage of the path matrix (0 = f64, 1 = f32); state_f32: per-path state stored as float (f32 paths only)
cudaError_t launch_step(int dtype, int state_f32, int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl,
                        int n_batch) {
    if (dtype == 1 && state_f32) return launch_step_f32s(degree, grid, a, s, pdl, n_batch);
    if (state_f32) return cudaErrorInvalidValue;
    return dtype == 1 ? launch_step_f32(degree, grid, a, s, pdl, n_batch) : launch_step_f64(degree, grid, a, s, pdl, n_batch);
}

// grid = SM count x resident blocks per SM: every block is co-resident, the grid-stride loop balances.
int step_grid_size(int dtype, int state_f32, int degree, int sm_count) {
    int nb = dtype == 1 ? (state_f32 ? step_occupancy_f32s(degree) : step_occupancy_f32(degree)) : step_occupancy_f64(degree);
    if (nb < 1) nb = 1;
    return sm_count * nb;
}

cudaError_t launch_sweep(int dtype, int state_f32, int degree, int lean, int grid, const SweepArgs& a, cudaStream_t s) {
    if (lean) return dtype == 1 ? launch_sweep_lean(state_f32, degree, grid, a, s) : cudaErrorInvalidValue;
    if (dtype == 1 && state_f32) return launch_sweep_f32s(degree, grid, a, s);
    if (state_f32) return cudaErrorInvalidValue;
    return dtype == 1 ? launch_sweep_f32(degree, grid, a, s) : launch_sweep_f64(degree, grid, a, s);
}

int sweep_grid_size(int dtype, int state_f32, int degree, int lean, int sm_count) {
    int nb;
    if (lean) nb = sweep_occupancy_lean(state_f32, degree);
    else nb = dtype == 1 ? (state_f32 ? sweep_occupancy_f32s(degree) : sweep_occupancy_f32(degree)) : sweep_occupancy_f64(degree);
    if (nb < 1) nb = 1;
    return sm_count * nb;                                  // cooperative launch: every block resident
}

cudaError_t launch_cluster_sweep(int dtype, int state_f32, int degree, const SweepArgs& a, cudaStream_t s) {
    if (dtype == 1 && state_f32) return launch_cluster_f32s(degree, a, s);
    if (state_f32) return cudaErrorInvalidValue;
    return dtype == 1 ? launch_cluster_f32(degree, a, s) : launch_cluster_f64(degree, a, s);
}

int64_t cluster_sweep_capacity(int dtype, int state_f32, int degree) {
    if (dtype == 1 && state_f32) return cluster_capacity_f32s(degree);
    if (state_f32) return 0;
    return dtype == 1 ? cluster_capacity_f32(degree) : cluster_capacity_f64(degree);
}

cudaError_t launch_solve(const SolveArgs& a, cudaStream_t s, bool pdl) {
    switch (a.spec.degree) {
        case 0: return launch_ex(lsm_solve_kernel<1>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 1: return launch_ex(lsm_solve_kernel<2>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 2: return launch_ex(lsm_solve_kernel<3>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 3: return launch_ex(lsm_solve_kernel<4>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 4: return launch_ex(lsm_solve_kernel<5>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 5: return launch_ex(lsm_solve_kernel<6>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 6: return launch_ex(lsm_solve_kernel<7>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 7: return launch_ex(lsm_solve_kernel<8>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 8: return launch_ex(lsm_solve_kernel<9>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 9: return launch_ex(lsm_solve_kernel<10>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
        case 10: return launch_ex(lsm_solve_kernel<11>, dim3(a.n_batch > 1 ? a.n_batch : 1), kSolveThreads, 0, s, pdl, a);
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
    const int grid = blocks_for(n, 256, device_sm_count() * 8);
    if (dtype == 1)
        continuation_kernel<float><<<grid, 256, 0, s>>>((const float*)x, n, gamma_dev, degree, mu, isg, clamp, out_dev);
    else
        continuation_kernel<double><<<grid, 256, 0, s>>>((const double*)x, n, gamma_dev, degree, mu, isg, clamp, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_discount(const double* cf_dev, const int64_t* tau_dev, int64_t n, int64_t t, double r, double dt,
                            double* y_dev, cudaStream_t s) {
    discount_kernel<<<blocks_for(n, 256, device_sm_count() * 8), 256, 0, s>>>(cf_dev, tau_dev, n, t, r, dt, y_dev);
    return cudaGetLastError();
}

cudaError_t launch_apply_exercise(double* cf_dev, int64_t* tau_dev, const double* ev_dev, const double* cont_dev,
                                  const int64_t* idx_dev, int64_t m, int64_t t, cudaStream_t s) {
    apply_exercise_kernel<<<blocks_for(m, 256, device_sm_count() * 8), 256, 0, s>>>(cf_dev, tau_dev, ev_dev, cont_dev, idx_dev, m, t);
    return cudaGetLastError();
}

cudaError_t launch_intrinsic(const double* S_dev, int64_t n, double K, int is_put, double* out_dev, cudaStream_t s) {
    intrinsic_kernel<<<blocks_for(n, 256, device_sm_count() * 8), 256, 0, s>>>(S_dev, n, K, is_put, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_basis_matrix(const double* X_dev, int64_t n, int basis, int degree, double* out_dev,
                                cudaStream_t s) {
    basis_matrix_kernel<<<blocks_for(n, 256, device_sm_count() * 8), 256, 0, s>>>(X_dev, n, basis, degree, out_dev);
    return cudaGetLastError();
}

}  // namespace amc
