// Per-step regression solve: (Hankel moment sums) -> coefficients of the continuation polynomial.
//
// Replaces, for one backward step, the O(P*k^2) part of the reference's
//   A = get_basis_polynomials(X[_scaled], basis, degree); coeffs = np.linalg.lstsq(A, Y, rcond=None)[0];
//   fitted = A @ coeffs                      (/root/reference/american_monte_carlo.py:110-122)
// by O(k^3) work on sums the streaming kernel has already reduced over ALL paths:
//   h[m] = sum_p z_p^m        m = 0..2d        z_p = (x_p - mu_ref) / sigma_ref
//   g[m] = sum_p z_p^m * y_p  m = 0..d
// The internal basis q_m(z) = z^m is well conditioned (z is roughly standardised); the user's basis
// (Power / Chebyshev / Legendre on raw or `scaling`-standardised x, no domain mapping --
// american_monte_carlo.py:98-106,111-114) enters only through the exact k x k change of basis M with
// A = Q M.  With G = Q^T Q = L L^T, the singular values of A are those of the k x k upper-triangular
// matrix B = L^T M, so numpy's rank rule  s_i > eps*max(P,k)*s_1  (lstsq rcond=None -> LAPACK gelsd)
// can be reproduced exactly, and the fitted values of the truncated solve are
//   fitted = Q L^-T U_r U_r^T L^-1 (Q^T y)
// where U_r are the left singular vectors of B above the cutoff.  When the rank is full this collapses
// to the normal-equation solution G gamma = Q^T y.  SURVEY.md Appendix A holds the NumPy derivation.
//
// Cost control: the SVD (one-sided Jacobi) is only needed to DECIDE the rank and, when the rank is not full,
// to project.  A cheap certificate -- s_min(B) >= 1/|B^-1|_F and s_max(B) <= |B|_F, with B^-1 by back
// substitution -- proves full rank for the vast majority of steps and skips the SVD; the answer is the same
// because a full-rank fit does not depend on the singular vectors.
//
// Everything is templated on K = degree+1 and fully unrolled so that the k x k matrices live in registers.
// The code is __host__ __device__: the same source runs (a) in the single-thread device solve between two
// streaming kernels and (b) in tests/native/solve_host.cpp, which lets the CPU-only test tier compare it with
// numpy.linalg.lstsq without a GPU.  The product library never calls it on the host.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define AMC_HD __host__ __device__ __forceinline__
#define AMC_UNROLL _Pragma("unroll")
#else
#define AMC_HD inline
#define AMC_UNROLL
#endif

namespace amc {

// 1 / sqrt(x): the device's rsqrt (<= 1 ulp, no division, no separate square root) -- the host build of this header (CPU
// test tier) has no such instruction and divides
AMC_HD double amc_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

constexpr int kMaxDegree = 10;           // notebook uses degree 10; plots.py sweeps 0..10
constexpr int kMaxK = kMaxDegree + 1;

enum BasisId { kPower = 0, kChebyshev = 1, kLegendre = 2, kLaguerre = 3 };

struct SolveSpec {
    int degree;               // d, k = d + 1
    int basis;                // BasisId
    int scaling;              // regression_estimate(..., scaling=True)
    int want_svd;             // always run the SVD and report singular values (diagnostics / tests)
    double scaling_factor;    // regression_estimate(..., scaling_factor=2)
    double n_paths;           // GLOBAL number of paths (all ranks) -- enters numpy's rcond = eps*max(P,k)
    int warp_solve;           // device only: try the warp-cooperative routine first (lsm_solve_warp.cuh)
    double inv_n_paths;       // 1 / n_paths, computed once on the host (0: not set, the solve divides itself): the division
                              // heads the dependent chain of every solve of a sweep
};

struct SolveResult {
    double gamma[kMaxK];      // fitted(x) = sum_m gamma[m] z^m,  z = (x - mu_ref)/sigma_ref
    double beta[kMaxK];       // numpy's lstsq coefficients in the USER basis (min-norm when truncated)
    double sv[kMaxK];         // singular values of the design matrix A, descending (only when the SVD ran)
    double mean_x, std_x;     // sample mean / population std of the column (np.mean, np.std)
    double pivot_loss;        // max over accepted Cholesky pivots of G_jj / pivot_j (>= 1): how many digits the
                              // factorisation of the internal Gram lost; ~cond(G).  Beyond ~1e12 the singular
                              // values it yields are no longer accurate to numpy's cutoff eps*max(P,k) and the rank
                              // decision (hence the fit) may deviate from lstsq -- reported, never hidden
    int rank;                 // numpy's reported rank
    int k_internal;           // internal monomials kept (== k unless the column is degenerate)
    int sweeps;               // Jacobi sweeps used; -1 = SVD skipped (full rank certified)
};

// Column j of M = coefficients (in z) of the user's basis polynomial phi_j(u), u = a + b z.
// Recurrences restate numpy.polynomial's definitions used at american_monte_carlo.py:99-101.
template <int K>
AMC_HD void build_change_of_basis(int basis, double a, double b, double (&M)[K][K]) {
    AMC_UNROLL
    for (int i = 0; i < K; ++i) {
        AMC_UNROLL
        for (int j = 0; j < K; ++j) M[i][j] = 0.0;
    }
    M[0][0] = 1.0;                                   // phi_0 = 1 for every family
    AMC_UNROLL
    for (int j = 1; j < K; ++j) {
        // u * phi_{j-1}: coefficient i is a*M[i][j-1] + b*M[i-1][j-1]
        AMC_UNROLL
        for (int i = 0; i <= j; ++i) {
            const double um = a * ((i < j) ? M[i][j - 1] : 0.0) + ((i > 0) ? b * M[i - 1][j - 1] : 0.0);
            const double p1 = (i < j) ? M[i][j - 1] : 0.0;
            const double p2 = (j >= 2 && i <= j - 2) ? M[i][j - 2] : 0.0;
            double v;
            switch (basis) {
                case kChebyshev:                     // T_1 = u, T_j = 2 u T_{j-1} - T_{j-2}
                    v = (j == 1) ? um : 2.0 * um - p2;
                    break;
                case kLegendre:                      // j P_j = (2j-1) u P_{j-1} - (j-1) P_{j-2}
                    v = ((2.0 * j - 1.0) * um - (j - 1.0) * p2) / (double)j;
                    break;
                case kLaguerre:                      // j L_j = (2j-1-u) L_{j-1} - (j-1) L_{j-2}
                    v = ((2.0 * j - 1.0) * p1 - um - (j - 1.0) * p2) / (double)j;
                    break;
                default:                             // Power: u^j
                    v = um;
            }
            M[i][j] = v;
        }
    }
}

// One-sided (Hestenes) Jacobi SVD of the K x K matrix B (overwritten by U*diag(s)); V accumulates the
// right rotations.  Columns are pre-sorted by norm (de Rijk) -- B's columns span ~20 orders of magnitude
// for unscaled high-degree bases and one-sided Jacobi keeps high RELATIVE accuracy under column scaling,
// which is what the rank rule needs.
template <int K>
AMC_HD int jacobi_svd(double (&B)[K][K], double (&V)[K][K], double (&s)[K]) {
    AMC_UNROLL
    for (int i = 0; i < K; ++i) {
        AMC_UNROLL
        for (int j = 0; j < K; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    }
    AMC_UNROLL
    for (int j = 0; j < K; ++j) {
        double nj = 0.0;
        AMC_UNROLL
        for (int i = 0; i < K; ++i) nj += B[i][j] * B[i][j];
        s[j] = nj;
    }
    // selection sort of the columns by decreasing norm (compare-and-swap network, static indices)
    AMC_UNROLL
    for (int j = 0; j < K - 1; ++j) {
        AMC_UNROLL
        for (int c = j + 1; c < K; ++c) {
            if (s[c] > s[j]) {
                AMC_UNROLL
                for (int i = 0; i < K; ++i) {
                    double t = B[i][j]; B[i][j] = B[i][c]; B[i][c] = t;
                    t = V[i][j]; V[i][j] = V[i][c]; V[i][c] = t;
                }
                double t = s[j]; s[j] = s[c]; s[c] = t;
            }
        }
    }
    const double tol2 = 1e-30;                       // (1e-15)^2: |a_p . a_q| <= 1e-15 |a_p| |a_q|
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        bool rotated = false;
        AMC_UNROLL
        for (int p = 0; p < K - 1; ++p) {
            AMC_UNROLL
            for (int q = p + 1; q < K; ++q) {
                double alpha = 0.0, beta = 0.0, gam = 0.0;
                AMC_UNROLL
                for (int i = 0; i < K; ++i) {
                    alpha += B[i][p] * B[i][p];
                    beta += B[i][q] * B[i][q];
                    gam += B[i][p] * B[i][q];
                }
                if (alpha > 0.0 && beta > 0.0 && gam * gam > tol2 * alpha * beta) {
                    rotated = true;
                    const double zeta = (beta - alpha) / (2.0 * gam);
                    const double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t);
                    const double sn = c * t;
                    AMC_UNROLL
                    for (int i = 0; i < K; ++i) {
                        const double bp = B[i][p], bq = B[i][q];
                        B[i][p] = c * bp - sn * bq;
                        B[i][q] = sn * bp + c * bq;
                        const double vp = V[i][p], vq = V[i][q];
                        V[i][p] = c * vp - sn * vq;
                        V[i][q] = sn * vp + c * vq;
                    }
                }
            }
        }
        if (!rotated) break;
    }
    AMC_UNROLL
    for (int j = 0; j < K; ++j) {
        double nj = 0.0;
        AMC_UNROLL
        for (int i = 0; i < K; ++i) nj += B[i][j] * B[i][j];
        s[j] = sqrt(nj);
    }
    return sweep;
}

// h: 2d+1 moment sums, g: d+1 cross sums, y_scale multiplies g (the step's growth factor
// exp(r*dt*t), because the kernels keep each path's cashflow discounted to time 0).
template <int K>
AMC_HD void lsm_solve_t(const SolveSpec& spec, const double* h, const double* g, double y_scale, double mu_ref,
                        double sigma_ref, SolveResult* out) {
    constexpr int D = K - 1;
    const double P = spec.n_paths;
    const double invP = (spec.inv_n_paths > 0.0) ? spec.inv_n_paths : 1.0 / P;
    AMC_UNROLL
    for (int i = 0; i < kMaxK; ++i) out->gamma[i] = out->beta[i] = out->sv[i] = 0.0;
    out->sweeps = 0;
    out->pivot_loss = 1.0;

    double Hn[2 * D + 1];
    AMC_UNROLL
    for (int m = 0; m <= 2 * D; ++m) Hn[m] = h[m] * invP;
    double b[K];
    AMC_UNROLL
    for (int i = 0; i < K; ++i) b[i] = g[i] * y_scale * invP;

    // column statistics (np.mean / np.std of x) from the first two z-moments
    const double mz = (D >= 1) ? Hn[(D >= 1) ? 1 : 0] : 0.0;
    if (D == 0) {
        out->mean_x = mu_ref;       // not observable from h[0] alone, and not needed: phi_0 = 1
        out->std_x = 0.0;
    } else {
        double vz = Hn[(D >= 1) ? 2 : 0] - mz * mz;
        if (vz < 0.0) vz = 0.0;
        out->mean_x = mu_ref + sigma_ref * mz;
        out->std_x = sigma_ref * sqrt(vz);
    }

    // Cholesky G = L L^T of the Hankel Gram, stopping at the first numerically dependent monomial.
    double L[K][K];
    AMC_UNROLL
    for (int i = 0; i < K; ++i) {
        AMC_UNROLL
        for (int j = 0; j < K; ++j) L[i][j] = 0.0;
    }
    double Linv[K];               // 1 / L[j][j]: one division per pivot, multiplications everywhere else
    AMC_UNROLL
    for (int i = 0; i < K; ++i) Linv[i] = 0.0;
    int kint = K;
    const double pivot_tol = 2e-14;
    // pivot loss = max_j G_jj / pivot_j, kept as the (numerator, denominator) pair that maximises it and divided once at
    // the end: the division is a diagnostic, the factorisation's dependent chain should not carry it
    double loss_num = 1.0, loss_den = 1.0;
    bool loss_overflow = false;
    AMC_UNROLL
    for (int j = 0; j < K; ++j) {
        if (kint == K) {
            double djj = Hn[2 * j];
            AMC_UNROLL
            for (int c = 0; c < j; ++c) djj -= L[j][c] * L[j][c];
            if (!(djj > pivot_tol * Hn[2 * j])) {
                kint = j;
                // A rejected pivot beyond the constant-column case (j >= 2) with more paths than monomials is not an
                // exactly degenerate column: the fit is truncated to the first j monomials WITHOUT numpy's singular-
                // value rule having been applied to A, so the loss is reported (the host warns from 1e12 on).
                if (j >= 2 && P > (double)K) {
                    if (djj > 0.0) {
                        if (Hn[2 * j] * loss_den > loss_num * djj) { loss_num = Hn[2 * j]; loss_den = djj; }
                    } else {
                        loss_overflow = true;
                    }
                }
            } else {
                // one reciprocal square root per pivot (L_jj = pivot * rsqrt(pivot)), multiplications everywhere else
                const double inv = amc_rsqrt(djj);
                const double ljj = djj * inv;
                if (Hn[2 * j] * loss_den > loss_num * djj) { loss_num = Hn[2 * j]; loss_den = djj; }
                L[j][j] = ljj;
                Linv[j] = inv;
                AMC_UNROLL
                for (int i = j + 1; i < K; ++i) {
                    double v = Hn[i + j];
                    AMC_UNROLL
                    for (int c = 0; c < j; ++c) v -= L[i][c] * L[j][c];
                    L[i][j] = v * inv;
                }
            }
        }
    }
    out->pivot_loss = loss_overflow ? 1e300 : loss_num / loss_den;
    out->k_internal = kint;
    if (kint == 0) { out->rank = 0; return; }         // no paths / all-NaN input

    // w = L^-1 b   (forward substitution on the kept block)
    double w[K];
    AMC_UNROLL
    for (int i = 0; i < K; ++i) {
        double v = b[i];
        AMC_UNROLL
        for (int c = 0; c < i; ++c) v -= L[i][c] * w[c];
        w[i] = (i < kint) ? v * Linv[i] : 0.0;
    }

    // change of basis for the user's polynomials: u = a + b z
    double ca, cb;
    if (spec.scaling) {
        const double sdev = out->std_x > 1e-6 ? out->std_x : 1e-6;       // max(np.std(X), 1e-6), amc.py:113
        const double den = spec.scaling_factor * sdev;
        ca = (mu_ref - out->mean_x) / den;
        cb = sigma_ref / den;
    } else {
        ca = mu_ref;
        cb = sigma_ref;
    }
    double M[K][K];
    build_change_of_basis<K>(spec.basis, ca, cb, M);

    double proj[K];
    const double sqrtP = sqrt(P);
    if (kint < K) {
        // Degenerate column (fewer than k distinct abscissae, e.g. t = 0 where every path sits at S0,
        // or P < k).  numpy's truncated SVD then projects y on the span of the surviving monomials:
        // with one distinct point the fit is mean(y).
        AMC_UNROLL
        for (int i = 0; i < K; ++i) proj[i] = w[i];
        out->rank = kint;
        if (kint == 1) {
            // A has identical rows a_j = phi_j(u0); min-norm beta = a * mean(y) / |a|^2, s_1 = sqrt(P)|a|
            double a2 = 0.0, arow[K];
            AMC_UNROLL
            for (int j = 0; j < K; ++j) {
                double v = 0.0, zp = 1.0;
                AMC_UNROLL
                for (int i = 0; i <= j; ++i) { v += M[i][j] * zp; zp *= mz; }
                arow[j] = v;
                a2 += v * v;
            }
            AMC_UNROLL
            for (int j = 0; j < K; ++j) out->beta[j] = arow[j] * b[0] / a2;
            out->sv[0] = sqrtP * sqrt(a2);
        } else {
            const double nanv = nan("");
            AMC_UNROLL
            for (int j = 0; j < K; ++j) out->beta[j] = nanv;
        }
    } else {
        // B = L^T M (upper triangular)
        double B[K][K];
        AMC_UNROLL
        for (int i = 0; i < K; ++i) {
            AMC_UNROLL
            for (int j = 0; j < K; ++j) {
                double v = 0.0;
                AMC_UNROLL
                for (int l = 0; l < K; ++l)
                    if (l >= i && l <= j) v += L[l][i] * M[l][j];
                B[i][j] = v;
            }
        }
        const double eps = 2.220446049250313e-16;
        const double rcond = eps * (P > (double)K ? P : (double)K);

        // full-rank certificate: s_min >= 1/|B^-1|_F  and  s_max <= |B|_F
        bool certified = false;
        if (!spec.want_svd) {
            double Bi[K][K];
            double nb = 0.0, nbi = 0.0;
            bool ok = true;
            AMC_UNROLL
            for (int j = 0; j < K; ++j) {
                AMC_UNROLL
                for (int i = 0; i < K; ++i) {
                    Bi[i][j] = 0.0;
                    if (i <= j) nb += B[i][j] * B[i][j];
                }
            }
            AMC_UNROLL
            for (int j = 0; j < K; ++j) {
                ok = ok && (B[j][j] != 0.0);
                Bi[j][j] = 1.0 / B[j][j];
            }
            AMC_UNROLL
            for (int j = 0; j < K; ++j) {
                AMC_UNROLL
                for (int i = K - 1; i >= 0; --i) {
                    if (i < j) {
                        double v = 0.0;
                        AMC_UNROLL
                        for (int l = 0; l < K; ++l)
                            if (l > i && l <= j) v += B[i][l] * Bi[l][j];
                        Bi[i][j] = -v * Bi[i][i];
                    }
                }
                AMC_UNROLL
                for (int i = 0; i < K; ++i)
                    if (i <= j) nbi += Bi[i][j] * Bi[i][j];
            }
            // 1/sqrt(nbi) > rcond*sqrt(nb) with a 0.1 % safety margin against rounding in the norms
            certified = ok && (nb == nb) && (nbi == nbi) && (1.0 > 1.002 * rcond * rcond * nb * nbi);
            if (certified) {
                out->rank = K;
                out->sweeps = -1;
                AMC_UNROLL
                for (int i = 0; i < K; ++i) {
                    proj[i] = w[i];
                    double v = 0.0;
                    AMC_UNROLL
                    for (int l = 0; l < K; ++l)
                        if (l >= i) v += Bi[i][l] * w[l];
                    out->beta[i] = v;                          // beta = B^-1 w
                }
            }
        }
        if (!certified) {
            double V[K][K], s[K];
            out->sweeps = jacobi_svd<K>(B, V, s);
            double smax = 0.0;
            AMC_UNROLL
            for (int j = 0; j < K; ++j) smax = s[j] > smax ? s[j] : smax;
            int rank = 0;
            AMC_UNROLL
            for (int i = 0; i < K; ++i) proj[i] = 0.0;
            AMC_UNROLL
            for (int j = 0; j < K; ++j) {
                if (s[j] > rcond * smax) {
                    ++rank;
                    double dot = 0.0;
                    AMC_UNROLL
                    for (int i = 0; i < K; ++i) dot += B[i][j] * w[i];       // (s_j u_j)^T w
                    const double c1 = dot / (s[j] * s[j]);                  // u_j^T w / s_j
                    AMC_UNROLL
                    for (int i = 0; i < K; ++i) {
                        proj[i] += B[i][j] * c1;                            // u_j (u_j^T w)
                        out->beta[i] += V[i][j] * c1;                       // v_j (u_j^T w) / s_j
                    }
                }
            }
            out->rank = rank;
            // singular values, descending (insertion into the output)
            AMC_UNROLL
            for (int j = 0; j < K; ++j) out->sv[j] = s[j] * sqrtP;
            AMC_UNROLL
            for (int a = 0; a < K - 1; ++a) {
                AMC_UNROLL
                for (int c = a + 1; c < K; ++c) {
                    if (out->sv[c] > out->sv[a]) { const double t = out->sv[a]; out->sv[a] = out->sv[c]; out->sv[c] = t; }
                }
            }
            if (rank == K) {
                AMC_UNROLL
                for (int i = 0; i < K; ++i) proj[i] = w[i];             // U U^T = I: skip the rounding
            }
        }
    }

    // gamma = L^-T proj   (back substitution on the kept block)
    AMC_UNROLL
    for (int i = K - 1; i >= 0; --i) {
        if (i < kint) {
            double v = proj[i];
            AMC_UNROLL
            for (int c = 0; c < K; ++c)
                if (c > i && c < kint) v -= L[c][i] * out->gamma[c];
            out->gamma[i] = v * Linv[i];
        }
    }
}

AMC_HD void lsm_solve(const SolveSpec& spec, const double* h, const double* g, double y_scale, double mu_ref,
                      double sigma_ref, SolveResult* out) {
    switch (spec.degree) {
        case 0: lsm_solve_t<1>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 1: lsm_solve_t<2>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 2: lsm_solve_t<3>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 3: lsm_solve_t<4>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 4: lsm_solve_t<5>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 5: lsm_solve_t<6>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 6: lsm_solve_t<7>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 7: lsm_solve_t<8>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 8: lsm_solve_t<9>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 9: lsm_solve_t<10>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        case 10: lsm_solve_t<11>(spec, h, g, y_scale, mu_ref, sigma_ref, out); break;
        default: out->rank = 0; out->k_internal = 0;
    }
}

}  // namespace amc
