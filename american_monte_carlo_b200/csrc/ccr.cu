// Counterparty-credit-risk exposures on the device: compute_ccr_exposures, /root/reference/american_monte_carlo.py
// lines 400-414 ("amc.py").  For every time step the reference takes the finite continuation values of all paths and
// reports (5th percentile, 95th percentile, mean) with numpy's default `linear` percentile rule.  Here the [P] vector
// is never materialised: its values are recomputed from the step's stored continuation polynomial (exactly as
// continuation_kernel does), and the four order statistics the two percentiles need are found by a radix select over
// the 64-bit order-preserving image of the doubles -- six histogram passes of 11+11+11+11+11+9 bits, each narrowing
// all four targets at once.  Integer histograms (shared-memory atomics, then global atomics) make the result exact
// and independent of the launch geometry; the mean is a fixed-order sum of per-block partials.
//   sel_hist_kernel : one pass over the values: per-target digit histogram of the keys that match the target's prefix
//   sel_scan_kernel : one block: finds each target's digit, narrows prefix and rank; after the last pass applies
//                     numpy's lerp (numpy/lib/_function_base_impl.py::_lerp) and writes (pfe_lo, pfe_hi, mean)
#include "common.cuh"
#include "kernels.h"

namespace amc {

__device__ __forceinline__ double ccr_value(const CcrSource& s, int64_t p) {
    if (s.zero) return 0.0;                                   // maturity: continuation values are zeros (amc.py:145)
    if (s.vals) return s.vals[p];
    const double x = s.x_f32 ? (double)static_cast<const float*>(s.x)[p] : static_cast<const double*>(s.x)[p];
    const double z = (x - s.mu) * s.isg;                      // same arithmetic as continuation_kernel
    double f = s.gam[s.degree];
    for (int m = s.degree - 1; m >= 0; --m) f = fma(f, z, s.gam[m]);
    return s.clamp ? ((f > 0.0 || f != f) ? f : 0.0) : f;
}

// order-preserving map double -> uint64 (finite values only)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ int pass_bits(int pass) { return pass < kSelPasses - 1 ? 11 : 9; }

__global__ void __launch_bounds__(256) sel_hist_kernel(const CcrSource s, int64_t n, const SelState* __restrict__ st,
                                                       int pass, unsigned long long* __restrict__ hist,
                                                       double* __restrict__ sum_partials) {
    __shared__ unsigned int sh[kSelTargets][kSelBins];
    __shared__ double red[8];
    for (int i = threadIdx.x; i < kSelTargets * kSelBins; i += blockDim.x) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const int done = st->bits_done;
    const int nb = pass_bits(pass);
    const int shift = 64 - done - nb;
    const unsigned long long mask = (1ull << nb) - 1ull;
    unsigned long long prefix[kSelTargets];
    bool live[kSelTargets];
#pragma unroll
    for (int t = 0; t < kSelTargets; ++t) {
        prefix[t] = st->prefix[t];
        // identical targets (e.g. all four before the first pass) share the histogram of the first of them
        live[t] = true;
        for (int u = 0; u < t; ++u) live[t] = live[t] && (prefix[u] != prefix[t]);
        if (done == 0) live[t] = (t == 0);
    }
    double acc[1] = {0.0};
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double v = ccr_value(s, p);
        if (!isfinite(v)) continue;                           // amc.py:404: np.isfinite filter
        if (pass == 0) acc[0] += v;
        const unsigned long long key = ordered_key(v);
        const unsigned long long hi = done ? (key >> (64 - done)) : 0ull;
        const unsigned int digit = (unsigned int)((key >> shift) & mask);
#pragma unroll
        for (int t = 0; t < kSelTargets; ++t)
            if (live[t] && hi == prefix[t]) atomicAdd(&sh[t][digit], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelTargets * kSelBins; i += blockDim.x) {
        const unsigned int c = (&sh[0][0])[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
    if (pass == 0) block_reduce_store<1, 256>(acc, red, sum_partials + blockIdx.x);
}

// <<<1, 32 * kSelTargets>>>: warp t narrows target t.
__global__ void sel_scan_kernel(SelState* __restrict__ st, unsigned long long* __restrict__ hist, int pass,
                                const double* __restrict__ sum_partials, int n_partials, double q_lo, double q_hi,
                                double* __restrict__ out3) {
    __shared__ unsigned long long s_prefix[kSelTargets], s_rank[kSelTargets];
    const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int done = st->bits_done;
    const int nb = pass_bits(pass);
    const int nbins = 1 << nb;
    if (pass == 0) {
        // number of finite values = total of the (shared) first histogram; ranks of the order statistics numpy's
        // linear rule needs: virtual index (N-1)q, its floor and floor+1 (both = N-1 when the index is the last one)
        __shared__ unsigned long long s_count;
        if (t == 0) {
            unsigned long long c = 0;
            for (int i = lane; i < nbins; i += 32) c += hist[i];
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) {
                s_count = c;
                double sum = 0.0;
                for (int i = 0; i < n_partials; ++i) sum += sum_partials[i];      // fixed order
                st->count = c;
                st->sum = sum;
                const double N = (double)c;
                const double q[2] = {q_lo, q_hi};
                for (int j = 0; j < 2; ++j) {
                    const double vi = (N - 1.0) * q[j];
                    double prev = floor(vi), next = prev + 1.0;
                    if (vi >= N - 1.0) { prev = N - 1.0; next = N - 1.0; }
                    if (vi < 0.0) { prev = 0.0; next = 0.0; }
                    st->rank[2 * j] = (unsigned long long)prev;
                    st->rank[2 * j + 1] = (unsigned long long)next;
                    st->frac[j] = vi - prev;
                }
            }
        }
        __syncthreads();
        if (s_count == 0) {                                    // amc.py:405-408: no finite value -> NaN
            if (threadIdx.x == 0) {
                const double nanv = __longlong_as_double(0x7ff8000000000000ll);
                out3[0] = nanv; out3[1] = nanv; out3[2] = nanv;
                st->bits_done = 64;
            }
            for (int i = threadIdx.x; i < kSelTargets * kSelBins; i += blockDim.x) hist[i] = 0ull;
            return;
        }
    }
    if (done >= 64) {                                          // empty column: nothing to narrow
        for (int i = threadIdx.x; i < kSelTargets * kSelBins; i += blockDim.x) hist[i] = 0ull;
        return;
    }
    __syncthreads();
    // which histogram does this target use (see sel_hist_kernel)
    int owner = t;
    if (done == 0) owner = 0;
    else
        for (int u = t - 1; u >= 0; --u)
            if (st->prefix[u] == st->prefix[t]) owner = u;
    const unsigned long long* h = hist + (size_t)owner * kSelBins;
    const unsigned long long rank = st->rank[t];
    // lane owns a contiguous run of bins; exclusive scan over lanes, then a serial walk inside the run
    const int per = nbins / 32;
    unsigned long long mine = 0;
    for (int i = 0; i < per; ++i) mine += h[lane * per + i];
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned long long excl = incl - mine;
    const bool has = rank >= excl && rank < incl;
    const unsigned int who = __ballot_sync(0xffffffffu, has);
    if (who != 0u && lane == __ffs(who) - 1) {
        unsigned long long below = excl;
        int digit = lane * per;
        for (int i = 0; i < per; ++i) {
            const unsigned long long c = h[lane * per + i];
            if (rank < below + c) { digit = lane * per + i; break; }
            below += c;
        }
        s_prefix[t] = (st->prefix[t] << nb) | (unsigned long long)digit;
        s_rank[t] = rank - below;
    }
    __syncthreads();
    if (threadIdx.x < kSelTargets) {
        st->prefix[threadIdx.x] = s_prefix[threadIdx.x];
        st->rank[threadIdx.x] = s_rank[threadIdx.x];
    }
    for (int i = threadIdx.x; i < kSelTargets * kSelBins; i += blockDim.x) hist[i] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
        st->bits_done = done + nb;
        if (done + nb == 64) {
            double r[2];
            for (int j = 0; j < 2; ++j) {
                const double a = key_value(s_prefix[2 * j]), b = key_value(s_prefix[2 * j + 1]);
                const double g = st->frac[j];
                const double diff = b - a;                     // numpy _lerp
                double v = a + diff * g;
                if (g >= 0.5) v = b - diff * (1.0 - g);
                r[j] = v;
            }
            out3[0] = r[0];
            out3[1] = r[1];
            out3[2] = st->sum / (double)st->count;
        }
    }
}

cudaError_t launch_ccr_hist(const CcrSource& s, int64_t n, const SelState* st, int pass, unsigned long long* hist,
                            double* sum_partials, int grid, cudaStream_t stream) {
    sel_hist_kernel<<<grid, 256, 0, stream>>>(s, n, st, pass, hist, sum_partials);
    return cudaGetLastError();
}

cudaError_t launch_ccr_scan(SelState* st, unsigned long long* hist, int pass, const double* sum_partials, int n_partials,
                            double q_lo, double q_hi, double* out3_dev, cudaStream_t stream) {
    sel_scan_kernel<<<1, 32 * kSelTargets, 0, stream>>>(st, hist, pass, sum_partials, n_partials, q_lo, q_hi, out3_dev);
    return cudaGetLastError();
}

}  // namespace amc
