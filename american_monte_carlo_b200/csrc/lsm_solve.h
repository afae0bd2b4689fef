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
// The code is __host__ __device__ so that the same source runs (a) in the single-warp device kernel
// between two streaming kernels and (b) in tests/native/solve_host.cpp, which lets the CPU-only test
// tier compare it with numpy.linalg.lstsq without a GPU.  The product library never calls it on the host.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define AMC_HD __host__ __device__ __forceinline__
#else
#define AMC_HD inline
#endif

namespace amc {

constexpr int kMaxDegree = 10;           // notebook uses degree 10; plots.py sweeps 0..10
constexpr int kMaxK = kMaxDegree + 1;

enum BasisId { kPower = 0, kChebyshev = 1, kLegendre = 2, kLaguerre = 3 };

struct SolveSpec {
    int degree;               // d, k = d + 1
    int basis;                // BasisId
    int scaling;              // regression_estimate(..., scaling=True)
    double scaling_factor;    // regression_estimate(..., scaling_factor=2)
    double n_paths;           // GLOBAL number of paths (all ranks) -- enters numpy's rcond = eps*max(P,k)
};

struct SolveResult {
    double gamma[kMaxK];      // fitted(x) = sum_m gamma[m] z^m,  z = (x - mu_ref)/sigma_ref
    double beta[kMaxK];       // numpy's lstsq coefficients in the USER basis (min-norm when truncated)
    double sv[kMaxK];         // singular values of the design matrix A, descending (numpy's 4th return)
    double mean_x, std_x;     // sample mean / population std of the column (np.mean, np.std)
    int rank;                 // numpy's reported rank
    int k_internal;           // internal monomials kept (== k unless the column is degenerate)
    int sweeps;               // Jacobi sweeps used (diagnostic)
};

// Multiply polynomial (coefficients in z, ascending, length n) by (a + b z) into out (length n+1).
AMC_HD void poly_mul_affine(const double* p, int n, double a, double b, double* out) {
    double carry = 0.0;
    for (int i = 0; i < n; ++i) {
        out[i] = a * p[i] + carry;
        carry = b * p[i];
    }
    out[n] = carry;
}

// Column j of M = coefficients (in z) of the user's basis polynomial phi_j(u), u = a + b z.
// Recurrences restate numpy.polynomial's definitions used at american_monte_carlo.py:99-101.
AMC_HD void build_change_of_basis(int basis, int k, double a, double b, double M[kMaxK][kMaxK]) {
    for (int i = 0; i < kMaxK; ++i)
        for (int j = 0; j < kMaxK; ++j) M[i][j] = 0.0;
    double prev2[kMaxK + 1], prev1[kMaxK + 1], cur[kMaxK + 1], tmp[kMaxK + 1];
    for (int i = 0; i <= kMaxK; ++i) prev2[i] = prev1[i] = cur[i] = tmp[i] = 0.0;
    prev1[0] = 1.0;                                  // phi_0 = 1 for every family
    M[0][0] = 1.0;
    for (int j = 1; j < k; ++j) {
        // tmp = u * phi_{j-1}
        poly_mul_affine(prev1, j, a, b, tmp);
        for (int i = 0; i <= j; ++i) {
            double v;
            switch (basis) {
                case kChebyshev:                     // T_1 = u, T_j = 2 u T_{j-1} - T_{j-2}
                    v = (j == 1) ? tmp[i] : 2.0 * tmp[i] - prev2[i];
                    break;
                case kLegendre:                      // j P_j = (2j-1) u P_{j-1} - (j-1) P_{j-2}
                    v = ((2.0 * j - 1.0) * tmp[i] - (j - 1.0) * prev2[i]) / (double)j;
                    break;
                case kLaguerre:                      // j L_j = (2j-1-u) L_{j-1} - (j-1) L_{j-2}
                    v = ((2.0 * j - 1.0) * prev1[i] - tmp[i] - (j - 1.0) * prev2[i]) / (double)j;
                    break;
                default:                             // Power: u^j
                    v = tmp[i];
            }
            cur[i] = v;
        }
        for (int i = 0; i <= j; ++i) {
            M[i][j] = cur[i];
            prev2[i] = prev1[i];
            prev1[i] = cur[i];
        }
        prev1[j] = cur[j];
        // prev2 must be phi_{j-1} padded with a zero at index j
        prev2[j] = 0.0;
    }
}

// One-sided (Hestenes) Jacobi SVD of the k x k matrix B (overwritten by U*diag(s)); V accumulates the
// right rotations.  Columns are pre-sorted by norm (de Rijk) -- B's columns span ~20 orders of
// magnitude for unscaled high-degree bases and one-sided Jacobi keeps high RELATIVE accuracy under
// column scaling, which is what the rank rule needs.
AMC_HD int jacobi_svd(int k, double B[kMaxK][kMaxK], double V[kMaxK][kMaxK], double s[kMaxK]) {
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    // sort columns by decreasing norm
    for (int j = 0; j < k; ++j) {
        double nj = 0.0;
        for (int i = 0; i < k; ++i) nj += B[i][j] * B[i][j];
        s[j] = nj;
    }
    for (int j = 0; j < k - 1; ++j) {
        int best = j;
        for (int c = j + 1; c < k; ++c)
            if (s[c] > s[best]) best = c;
        if (best != j) {
            for (int i = 0; i < k; ++i) {
                double t = B[i][j]; B[i][j] = B[i][best]; B[i][best] = t;
                t = V[i][j]; V[i][j] = V[i][best]; V[i][best] = t;
            }
            double t = s[j]; s[j] = s[best]; s[best] = t;
        }
    }
    const double tol = 1e-15;
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < k - 1; ++p) {
            for (int q = p + 1; q < k; ++q) {
                double alpha = 0.0, beta = 0.0, gam = 0.0;
                for (int i = 0; i < k; ++i) {
                    alpha += B[i][p] * B[i][p];
                    beta += B[i][q] * B[i][q];
                    gam += B[i][p] * B[i][q];
                }
                if (alpha == 0.0 || beta == 0.0) continue;
                if (fabs(gam) <= tol * sqrt(alpha) * sqrt(beta)) continue;
                rotated = true;
                double zeta = (beta - alpha) / (2.0 * gam);
                double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t);
                double sn = c * t;
                for (int i = 0; i < k; ++i) {
                    double bp = B[i][p], bq = B[i][q];
                    B[i][p] = c * bp - sn * bq;
                    B[i][q] = sn * bp + c * bq;
                    double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - sn * vq;
                    V[i][q] = sn * vp + c * vq;
                }
            }
        }
        if (!rotated) break;
    }
    for (int j = 0; j < k; ++j) {
        double nj = 0.0;
        for (int i = 0; i < k; ++i) nj += B[i][j] * B[i][j];
        s[j] = sqrt(nj);
    }
    return sweep;
}

// h: 2d+1 moment sums, g: d+1 cross sums, y_scale multiplies g (the step's growth factor
// exp(r*dt*t), because the kernels keep each path's cashflow discounted to time 0).
AMC_HD void lsm_solve(const SolveSpec& spec, const double* h, const double* g, double y_scale,
                      double mu_ref, double sigma_ref, SolveResult* out) {
    const int d = spec.degree;
    const int k = d + 1;
    const double P = spec.n_paths;
    const double invP = 1.0 / P;
    for (int i = 0; i < kMaxK; ++i) out->gamma[i] = out->beta[i] = out->sv[i] = 0.0;
    out->sweeps = 0;

    double Hn[2 * kMaxK];
    for (int m = 0; m <= 2 * d; ++m) Hn[m] = h[m] * invP;
    double b[kMaxK];
    for (int i = 0; i < k; ++i) b[i] = g[i] * y_scale * invP;

    // column statistics (np.mean / np.std of x) from the first two z-moments
    const double mz = (d >= 1) ? Hn[1] : 0.0;
    double vz = (d >= 1) ? Hn[2] - mz * mz : 0.0;
    if (d == 0) { out->mean_x = mu_ref; out->std_x = 0.0; }   // not observable from h[0] alone, not needed
    else {
        if (vz < 0.0) vz = 0.0;
        out->mean_x = mu_ref + sigma_ref * mz;
        out->std_x = sigma_ref * sqrt(vz);
    }

    // Cholesky G = L L^T of the Hankel Gram, stopping at the first numerically dependent monomial.
    double L[kMaxK][kMaxK];
    for (int i = 0; i < kMaxK; ++i)
        for (int j = 0; j < kMaxK; ++j) L[i][j] = 0.0;
    int kint = k;
    const double pivot_tol = 2e-14;
    for (int j = 0; j < k; ++j) {
        double djj = Hn[2 * j];
        for (int c = 0; c < j; ++c) djj -= L[j][c] * L[j][c];
        if (!(djj > pivot_tol * Hn[2 * j])) { kint = j; break; }
        const double ljj = sqrt(djj);
        L[j][j] = ljj;
        for (int i = j + 1; i < k; ++i) {
            double v = Hn[i + j];
            for (int c = 0; c < j; ++c) v -= L[i][c] * L[j][c];
            L[i][j] = v / ljj;
        }
    }
    out->k_internal = kint;
    if (kint == 0) { out->rank = 0; return; }         // no paths / all-NaN input

    // w = L^-1 b   (forward substitution on the kept block)
    double w[kMaxK];
    for (int i = 0; i < kint; ++i) {
        double v = b[i];
        for (int c = 0; c < i; ++c) v -= L[i][c] * w[c];
        w[i] = v / L[i][i];
    }

    // change of basis for the user's polynomials: u = a + b z
    double ca, cb;
    if (spec.scaling) {
        double sdev = out->std_x > 1e-6 ? out->std_x : 1e-6;       // max(np.std(X), 1e-6), amc.py:113
        double den = spec.scaling_factor * sdev;
        ca = (mu_ref - out->mean_x) / den;
        cb = sigma_ref / den;
    } else {
        ca = mu_ref;
        cb = sigma_ref;
    }
    double M[kMaxK][kMaxK];
    build_change_of_basis(spec.basis, k, ca, cb, M);

    double proj[kMaxK];
    const double sqrtP = sqrt(P);
    if (kint < k) {
        // Degenerate column (fewer than k distinct abscissae, e.g. t = 0 where every path sits at S0,
        // or P < k).  numpy's truncated SVD then projects y on the span of the surviving monomials:
        // with one distinct point the fit is mean(y).
        for (int i = 0; i < kint; ++i) proj[i] = w[i];
        out->rank = kint;
        if (kint == 1) {
            // A has identical rows a_j = phi_j(u0); min-norm beta = a * mean(y) / |a|^2, s_1 = sqrt(P)|a|
            double a2 = 0.0, arow[kMaxK];
            for (int j = 0; j < k; ++j) {
                double v = 0.0, zp = 1.0;
                for (int i = 0; i <= j; ++i) { v += M[i][j] * zp; zp *= mz; }
                arow[j] = v; a2 += v * v;
            }
            for (int j = 0; j < k; ++j) out->beta[j] = arow[j] * b[0] / a2;
            out->sv[0] = sqrtP * sqrt(a2);
        } else {
            const double nanv = nan("");
            for (int j = 0; j < k; ++j) out->beta[j] = nanv;
        }
    } else {
        // B = L^T M (upper triangular), SVD by one-sided Jacobi
        double B[kMaxK][kMaxK], V[kMaxK][kMaxK], s[kMaxK];
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                double v = 0.0;
                for (int l = i; l <= j; ++l) v += L[l][i] * M[l][j];
                B[i][j] = v;
            }
        out->sweeps = jacobi_svd(k, B, V, s);
        // order singular values descending (indices)
        int ord[kMaxK];
        for (int j = 0; j < k; ++j) ord[j] = j;
        for (int j = 0; j < k - 1; ++j) {
            int best = j;
            for (int c = j + 1; c < k; ++c)
                if (s[ord[c]] > s[ord[best]]) best = c;
            int t = ord[j]; ord[j] = ord[best]; ord[best] = t;
        }
        const double smax = s[ord[0]];
        const double eps = 2.220446049250313e-16;
        const double rcond = eps * (P > (double)k ? P : (double)k);
        int rank = 0;
        for (int j = 0; j < k; ++j) {
            out->sv[j] = s[ord[j]] * sqrtP;
            if (s[ord[j]] > rcond * smax) ++rank;
        }
        out->rank = rank;
        for (int i = 0; i < k; ++i) { proj[i] = 0.0; }
        for (int jj = 0; jj < rank; ++jj) {
            const int j = ord[jj];
            double dot = 0.0;
            for (int i = 0; i < k; ++i) dot += B[i][j] * w[i];       // (s_j u_j)^T w
            const double c1 = dot / (s[j] * s[j]);                  // u_j^T w / s_j
            for (int i = 0; i < k; ++i) {
                proj[i] += B[i][j] * c1;                            // u_j (u_j^T w)
                out->beta[i] += V[i][j] * c1;                       // v_j (u_j^T w) / s_j
            }
        }
        if (rank == k)
            for (int i = 0; i < k; ++i) proj[i] = w[i];             // U U^T = I: skip the rounding
    }

    // gamma = L^-T proj   (back substitution on the kept block)
    for (int i = kint - 1; i >= 0; --i) {
        double v = proj[i];
        for (int c = i + 1; c < kint; ++c) v -= L[c][i] * out->gamma[c];
        out->gamma[i] = v / L[i][i];
    }
}

}  // namespace amc
