// Warp-cooperative version of the per-step regression solve: full internal rank (no rejected Cholesky pivot), then
// either the certified full-rank fit or -- round 2 -- the rank-truncated fit through a warp-parallel Jacobi SVD (see
// lsm_solve.h for the method and the scalar routine that handles everything else: degenerate columns, want_svd
// diagnostics).
//
// One warp, lane i owns row / column i of the k x k matrices, which live in shared memory.  Every element is computed
// by exactly the expression (and operation order) the scalar routine uses, so the two give the same bits; only the two
// Frobenius norms of the certificate are summed in a different order, which can matter for the certificate's yes/no
// within rounding of its 0.2 % margin and never for the coefficients.  The point is latency: the scalar solve keeps its
// matrices in registers up to k = 6 and spills beyond (19 us per step at degree 8); here the dependent chain is the
// Cholesky's k square-root/reciprocal pairs plus O(k^2) multiply-adds per lane (~3 us at degree 8).
#pragma once
#include "lsm_solve.h"

namespace amc {

template <int K>
struct SolveShared {
    double Hn[2 * K];
    double L[K][K + 1], M[K][K + 1], B[K][K + 1], Bi[K][K + 1];
    double w[K], Linv[K], dinv[K], gamma[K], beta[K];
};

// Returns true when the step was solved here (outputs written by lane 0); false -> run the scalar routine.
template <int K>
__device__ bool lsm_solve_warp(const SolveSpec& spec, const double* hsum /* [2d]: sum z^m, m = 1..2d */,
                               const double* gsum /* [d+1]: sum z^m y */, double y_scale, double mu_ref, double sigma_ref,
                               SolveShared<K>& sh, double* gamma_out, double* beta_out, double* sv_out,
                               double* mean_std_out, int* rank_out) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int D = K - 1;
    const int lane = threadIdx.x & 31;
    if (spec.want_svd || D == 0 || !spec.warp_solve) return false;
    const double P = spec.n_paths;
    const double invP = (spec.inv_n_paths > 0.0) ? spec.inv_n_paths : 1.0 / P;
    for (int m = lane; m <= 2 * D; m += 32) sh.Hn[m] = (m == 0 ? P : hsum[m - 1]) * invP;
    const double bi = (lane < K) ? gsum[lane] * y_scale * invP : 0.0;
    __syncwarp();

    const double mz = sh.Hn[1];
    double vz = sh.Hn[2] - mz * mz;
    if (vz < 0.0) vz = 0.0;
    const double mean_x = mu_ref + sigma_ref * mz;
    const double std_x = sigma_ref * sqrt(vz);

    // Cholesky G = L L^T of the Hankel Gram (lsm_solve.h, same expressions)
    for (int i = lane; i < K * (K + 1); i += 32) (&sh.L[0][0])[i] = 0.0;
    __syncwarp();
    const double pivot_tol = 2e-14;
    double pivot_loss = 1.0;
    for (int j = 0; j < K; ++j) {
        double inv = 0.0;
        int ok = 1;
        if (lane == j) {
            double djj = sh.Hn[2 * j];
            for (int c = 0; c < j; ++c) djj -= sh.L[j][c] * sh.L[j][c];
            if (!(djj > pivot_tol * sh.Hn[2 * j])) {
                ok = 0;
            } else {
                inv = amc_rsqrt(djj);                       // as the scalar routine: L_jj = pivot * rsqrt(pivot)
                const double ljj = djj * inv;
                pivot_loss = sh.Hn[2 * j] / djj;            // per-lane, off the other lanes' critical path
                sh.L[j][j] = ljj;
                sh.Linv[j] = inv;
            }
        }
        ok = __shfl_sync(FULL, ok, j);
        if (!ok) return false;                               // numerically dependent monomial: scalar routine
        inv = __shfl_sync(FULL, inv, j);
        if (lane > j && lane < K) {
            double v = sh.Hn[lane + j];
            for (int c = 0; c < j; ++c) v -= sh.L[lane][c] * sh.L[j][c];
            sh.L[lane][j] = v * inv;
        }
        __syncwarp();
    }

    // w = L^-1 b: lane i carries v_i; after w_c is known every later row subtracts L[i][c] w_c (ascending c, as scalar)
    double v = bi;
    for (int c = 0; c < K; ++c) {
        double wc = (lane == c) ? v * sh.Linv[c] : 0.0;
        wc = __shfl_sync(FULL, wc, c);
        if (lane == c) sh.w[c] = wc;
        if (lane > c && lane < K) v -= sh.L[lane][c] * wc;
    }

    // change of basis for the user's polynomials, u = ca + cb z (lsm_solve.h build_change_of_basis, column by column)
    double ca, cb;
    if (spec.scaling) {
        const double sdev = std_x > 1e-6 ? std_x : 1e-6;
        const double den = spec.scaling_factor * sdev;
        ca = (mu_ref - mean_x) / den;
        cb = sigma_ref / den;
    } else {
        ca = mu_ref;
        cb = sigma_ref;
    }
    for (int i = lane; i < K * (K + 1); i += 32) (&sh.M[0][0])[i] = 0.0;
    __syncwarp();
    if (lane == 0) sh.M[0][0] = 1.0;
    __syncwarp();
    for (int j = 1; j < K; ++j) {
        if (lane <= j) {
            const int i = lane;
            const double um = ca * ((i < j) ? sh.M[i][j - 1] : 0.0) + ((i > 0) ? cb * sh.M[i - 1][j - 1] : 0.0);
            const double p1 = (i < j) ? sh.M[i][j - 1] : 0.0;
            const double p2 = (j >= 2 && i <= j - 2) ? sh.M[i][j - 2] : 0.0;
            double val;
            switch (spec.basis) {
                case kChebyshev: val = (j == 1) ? um : 2.0 * um - p2; break;
                case kLegendre: val = ((2.0 * j - 1.0) * um - (j - 1.0) * p2) / (double)j; break;
                case kLaguerre: val = ((2.0 * j - 1.0) * p1 - um - (j - 1.0) * p2) / (double)j; break;
                default: val = um;
            }
            sh.M[i][j] = val;
        }
        __syncwarp();
    }

    // B = L^T M (upper triangular): lane j forms column j
    if (lane < K) {
        const int j = lane;
        for (int i = 0; i < K; ++i) {
            double s = 0.0;
            for (int l = 0; l < K; ++l)
                if (l >= i && l <= j) s += sh.L[l][i] * sh.M[l][j];
            sh.B[i][j] = s;
        }
    }
    __syncwarp();

    // full-rank certificate: s_min >= 1/|B^-1|_F, s_max <= |B|_F; lane j inverts column j by back substitution
    const double eps = 2.220446049250313e-16;
    const double rcond = eps * (P > (double)K ? P : (double)K);
    int okd = 1;
    if (lane < K) {
        okd = (sh.B[lane][lane] != 0.0);
        sh.dinv[lane] = 1.0 / sh.B[lane][lane];
    }
    okd = __all_sync(FULL, okd);
    __syncwarp();
    double nb = 0.0, nbi = 0.0;
    if (lane < K) {
        const int j = lane;
        for (int i = 0; i < K; ++i) sh.Bi[i][j] = 0.0;
        sh.Bi[j][j] = sh.dinv[j];
        for (int i = K - 1; i >= 0; --i) {
            if (i < j) {
                double s = 0.0;
                for (int l = 0; l < K; ++l)
                    if (l > i && l <= j) s += sh.B[i][l] * sh.Bi[l][j];
                sh.Bi[i][j] = -s * sh.dinv[i];
            }
        }
        for (int i = 0; i <= j; ++i) {
            nb += sh.B[i][j] * sh.B[i][j];
            nbi += sh.Bi[i][j] * sh.Bi[i][j];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nb += __shfl_xor_sync(FULL, nb, o);
        nbi += __shfl_xor_sync(FULL, nbi, o);
    }
    const bool certified = okd && (nb == nb) && (nbi == nbi) && (1.0 > 1.002 * rcond * rcond * nb * nbi);
    if (!certified && !((nb == nb) && nb > 0.0)) return false;      // NaN / empty design: scalar routine
    __syncwarp();

    int rank = K;
    if (certified) {
        // beta = B^-1 w (numpy's coefficients in the user's basis); the projection is w itself
        if (lane < K) {
            double s = 0.0;
            for (int l = 0; l < K; ++l)
                if (l >= lane) s += sh.Bi[lane][l] * sh.w[l];
            sh.beta[lane] = s;
            sh.dinv[lane] = sh.w[lane];                              // proj
        }
    } else {
        // Rank decision by the singular values of B = L^T M: one-sided (Hestenes) Jacobi, the rotations of one round of a
        // round-robin tournament done by different lanes at the same time (the pairs of a round are disjoint columns).
        // Same convergence test and rank rule as the scalar routine (lsm_solve.h jacobi_svd); V (in the storage of Bi)
        // accumulates the right rotations for numpy's minimum-norm coefficients.  At k >= 7 the scalar routine spills its
        // matrices and takes 100-300 us per step; this takes a few rounds of ~1 us.
        // columns pre-sorted by decreasing norm (de Rijk), as in the scalar routine: B's columns span ~20 orders of
        // magnitude for unscaled high-degree bases and the sorted order is what makes Jacobi converge in 2-3 sweeps.
        // Lane j finds the position of its column, moves it there through the (now free) storage of M; V starts as the
        // same permutation.
        for (int i = lane; i < K * (K + 1); i += 32) (&sh.Bi[0][0])[i] = 0.0;
        double nj = 0.0;
        if (lane < K)
            for (int i = 0; i < K; ++i) nj += sh.B[i][lane] * sh.B[i][lane];
        int pos = 0;
        for (int j = 0; j < K; ++j) {
            const double no = __shfl_sync(FULL, nj, j);
            if (lane < K && (no > nj || (no == nj && j < lane))) ++pos;
        }
        __syncwarp();
        if (lane < K) {
            for (int i = 0; i < K; ++i) sh.M[i][pos] = sh.B[i][lane];
            sh.Bi[lane][pos] = 1.0;
        }
        __syncwarp();
        if (lane < K)
            for (int i = 0; i < K; ++i) sh.B[i][lane] = sh.M[i][lane];
        __syncwarp();
        constexpr int N2 = (K + 1) / 2;              // pairs per round
        constexpr int MP = 2 * N2;                   // players: k columns (+ a bye when k is odd)
        const double tol2 = 1e-30;                   // (1e-15)^2: |a_p . a_q| <= 1e-15 |a_p| |a_q|
        for (int sweep = 0; sweep < 60; ++sweep) {
            int rotated = 0;
            for (int r = 0; r < MP - 1; ++r) {
                if (lane < N2) {
                    int p, q;
                    if (lane == 0) { p = r; q = MP - 1; }
                    else { p = (r + lane) % (MP - 1); q = (r + (MP - 1) - lane) % (MP - 1); }
                    if (p > q) { const int t = p; p = q; q = t; }
                    if (q < K) {
                        // both columns in registers: one read, one write of each per rotation
                        double bp[K], bq[K];
                        double alpha = 0.0, beta = 0.0, gam = 0.0;
#pragma unroll
                        for (int i = 0; i < K; ++i) {
                            bp[i] = sh.B[i][p];
                            bq[i] = sh.B[i][q];
                        }
#pragma unroll
                        for (int i = 0; i < K; ++i) {
                            alpha = fma(bp[i], bp[i], alpha);
                            beta = fma(bq[i], bq[i], beta);
                            gam = fma(bp[i], bq[i], gam);
                        }
                        if (alpha > 0.0 && beta > 0.0 && gam * gam > tol2 * alpha * beta) {
                            rotated = 1;
                            const double zeta = (beta - alpha) * (0.5 / gam);
                            const double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                            const double c = rsqrt(1.0 + t * t);
                            const double sn = c * t;
#pragma unroll
                            for (int i = 0; i < K; ++i) {
                                sh.B[i][p] = c * bp[i] - sn * bq[i];
                                sh.B[i][q] = sn * bp[i] + c * bq[i];
                            }
#pragma unroll
                            for (int i = 0; i < K; ++i) {
                                const double vp = sh.Bi[i][p], vq = sh.Bi[i][q];
                                sh.Bi[i][p] = c * vp - sn * vq;
                                sh.Bi[i][q] = sn * vp + c * vq;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (!__any_sync(FULL, rotated)) break;
        }
        // singular values = column norms; rank rule of numpy.linalg.lstsq(rcond=None) with the GLOBAL path count
        double sj = 0.0;
        if (lane < K) {
            for (int i = 0; i < K; ++i) sj += sh.B[i][lane] * sh.B[i][lane];
            sj = sqrt(sj);
        }
        double smax = sj;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) smax = fmax(smax, __shfl_xor_sync(FULL, smax, o));
        const bool kept = (lane < K) && (sj > rcond * smax);
        rank = __popc(__ballot_sync(FULL, kept));
        // c1_j = (s_j u_j)^T w / s_j^2 for the kept directions (0 otherwise)
        double c1 = 0.0;
        if (kept) {
            double dot = 0.0;
            for (int i = 0; i < K; ++i) dot += sh.B[i][lane] * sh.w[i];
            c1 = dot / (sj * sj);
        }
        if (lane < K) {
            sh.gamma[lane] = c1;                                       // scratch: read by every lane below
            sh.beta[lane] = 0.0;
        }
        __syncwarp();
        double proj = 0.0, bet = 0.0;
        if (lane < K) {
            for (int j = 0; j < K; ++j) {
                proj += sh.B[lane][j] * sh.gamma[j];                   // u_j (u_j^T w)
                bet += sh.Bi[lane][j] * sh.gamma[j];                   // v_j (u_j^T w) / s_j
            }
            if (rank == K) proj = sh.w[lane];                          // U U^T = I: skip the rounding
        }
        __syncwarp();
        if (lane < K) {
            sh.dinv[lane] = proj;
            sh.beta[lane] = bet;
            sh.Hn[lane] = sj;                                          // Hn is free by now: singular values, unsorted
        }
    }
    __syncwarp();
    // gamma = L^-T proj by back substitution (lane 0, scalar order)
    if (lane == 0) {
        for (int i = K - 1; i >= 0; --i) {
            double s = sh.dinv[i];
            for (int c = 0; c < K; ++c)
                if (c > i) s -= sh.L[c][i] * sh.gamma[c];
            sh.gamma[i] = s * sh.Linv[i];
        }
        if (!certified && sv_out) {
            // singular values of the design matrix, descending
            double sv[K];
            for (int j = 0; j < K; ++j) sv[j] = sh.Hn[j] * sqrt(P);
            for (int a = 0; a < K - 1; ++a)
                for (int c = a + 1; c < K; ++c)
                    if (sv[c] > sv[a]) { const double t = sv[a]; sv[a] = sv[c]; sv[c] = t; }
            for (int j = 0; j < kMaxK; ++j) sv_out[j] = j < K ? sv[j] : 0.0;
        }
    }
    __syncwarp();
    if (lane < kMaxK) {
        gamma_out[lane] = lane < K ? sh.gamma[lane] : 0.0;
        if (beta_out) beta_out[lane] = lane < K ? sh.beta[lane] : 0.0;
        if (sv_out && certified) sv_out[lane] = 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pivot_loss = fmax(pivot_loss, __shfl_xor_sync(FULL, pivot_loss, o));
    if (lane == 0) {
        if (mean_std_out) { mean_std_out[0] = mean_x; mean_std_out[1] = std_x; mean_std_out[2] = pivot_loss; }
        if (rank_out) rank_out[0] = rank;
    }
    return true;
}

}  // namespace amc
