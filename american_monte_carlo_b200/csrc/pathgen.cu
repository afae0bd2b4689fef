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
#include "common.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace amc {

// ---------------------------------------------------------------------------------------------------------
// K1, f32 storage: 4 adjacent paths per thread, 4 steps per Philox call and path.  Everything stays on the FP32 /
// integer / MUFU pipes (no f32<->f64 conversions, no FP64 adds): the cumulative log-price is a Kahan-compensated float
// sum kept in log2 units, so each step costs one FFMA + four FADD for the sum and one MUFU.EX2 + one FMUL for the
// price.  Accuracy: the compensated sum is good to ~1e-8 in the exponent, below the 6e-8 rounding of the float the
// price is stored in; the float-rounded step constants (drift, vol) are off by < 3e-8 relative, ~1e-8 on the price.
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

__global__ void __launch_bounds__(256) philox_paths_f32_kernel(float* __restrict__ S, int64_t ld, int n_steps,
                                                               int64_t n_local, int64_t path_offset, GbmParams g,
                                                               uint32_t k0, uint32_t k1) {
    const int64_t n_vec = (n_local + 3) >> 2;
    const float S0f = (float)g.S0;
    const float d2 = (float)(g.drift * 1.4426950408889634), v2 = (float)(g.vol * 1.4426950408889634);   // log2 units
    const float neg2ln2 = -1.3862943611198906f;             // -2 ln u = (-2 ln 2) log2 u
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p0 = v << 2;
        float L[4] = {0.f, 0.f, 0.f, 0.f}, C[4] = {0.f, 0.f, 0.f, 0.f};
        float* out_col = S + p0;
        st_stream(reinterpret_cast<float4*>(out_col), make_float4(S0f, S0f, S0f, S0f));
        for (int t0 = 0; t0 < n_steps; t0 += 4) {
            float z[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t gid = (uint64_t)(path_offset + p0 + i);
                const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(t0 >> 2),
                                                kPhiloxDomain, k0, k1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // u1 in (0, 1]: full 32-bit resolution in the tail (small integers convert exactly)
                    const float u1 = fmaf((float)r.v[2 * h], 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                    const float rad = sqrt_approx(neg2ln2 * lg2_approx(u1));
                    // angle: the low 23 bits become the mantissa of a float in [1, 2) (one LOP3, no int->float
                    // conversion on the MUFU pipe); 2 pi f - 3 pi lies in [-pi, pi)
                    const float f12 = __uint_as_float((r.v[2 * h + 1] & 0x007fffffu) | 0x3f800000u);
                    const float ang = fmaf(f12, 6.283185307179586f, -9.42477796076938f);
                    z[i][2 * h] = rad * cos_approx(ang);
                    z[i][2 * h + 1] = rad * sin_approx(ang);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (t0 + j < n_steps) {
                    out_col += ld;
                    float out[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float y = fmaf(v2, z[i][j], d2) - C[i];        // Kahan: carry the rounding of the sum
                        const float t = L[i] + y;
                        C[i] = (t - L[i]) - y;
                        L[i] = t;
                        out[i] = S0f * ex2_approx(t);
                    }
                    st_stream(reinterpret_cast<float4*>(out_col), make_float4(out[0], out[1], out[2], out[3]));
                }
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

cudaError_t launch_generate_philox(int dtype, void* S, int64_t ld, int n_steps, int64_t n_local, int64_t path_offset,
                                   GbmParams g, uint64_t seed, int sm_count, cudaStream_t s) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int64_t n_vec = dtype == 1 ? (n_local + 3) / 4 : (n_local + 1) / 2;
    int64_t blocks = (n_vec + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (dtype == 1)
        philox_paths_f32_kernel<<<(int)blocks, 256, 0, s>>>((float*)S, ld, n_steps, n_local, path_offset, g, k0, k1);
    else
        philox_paths_f64_kernel<<<(int)blocks, 256, 0, s>>>((double*)S, ld, n_steps, n_local, path_offset, g, k0, k1);
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
    if (blocks > 148 * 32) blocks = 148 * 32;
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
    if (tiles > 148 * 32) tiles = 148 * 32;
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
    if (blocks > 148 * 16) blocks = 148 * 16;
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
    if (blocks > 148 * 16) blocks = 148 * 16;
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
    if (blocks > 148 * 16) blocks = 148 * 16;
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
    if (blocks > 148 * 8) blocks = 148 * 8;
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
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    hit_matrix_kernel<<<(int)blocks, 256, 0, s>>>(first_hit_dev, n_cols, n_local, out_dev);
    return cudaGetLastError();
}

}  // namespace amc
