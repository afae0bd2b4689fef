// Host build of american_monte_carlo_b200/csrc/lsm_solve.h for the CPU-only test tier.
// TEST INFRASTRUCTURE: compiled by tests/conftest.py into tests/native/_build/, never shipped,
// never linked into libamc.so.  It lets `pytest -m "not gpu"` compare the device solver's source
// with numpy.linalg.lstsq without a GPU.
#include "../../american_monte_carlo_b200/csrc/lsm_solve.h"

extern "C" int amc_test_lsm_solve(int degree, int basis, int scaling, int want_svd, double scaling_factor, double n_paths,
                                  const double* h, const double* g, double y_scale, double mu_ref,
                                  double sigma_ref, double* gamma, double* beta, double* sv, double* stats,
                                  int* info) {
    if (degree < 0 || degree > amc::kMaxDegree) return 1;
    amc::SolveSpec spec;
    spec.degree = degree;
    spec.basis = basis;
    spec.scaling = scaling;
    spec.want_svd = want_svd;
    spec.scaling_factor = scaling_factor;
    spec.n_paths = n_paths;
    spec.warp_solve = 0;
    spec.inv_n_paths = n_paths > 0.0 ? 1.0 / n_paths : 0.0;
    amc::SolveResult res;
    amc::lsm_solve(spec, h, g, y_scale, mu_ref, sigma_ref, &res);
    for (int i = 0; i <= degree; ++i) { gamma[i] = res.gamma[i]; beta[i] = res.beta[i]; sv[i] = res.sv[i]; }
    stats[0] = res.mean_x; stats[1] = res.std_x;
    info[0] = res.rank; info[1] = res.k_internal; info[2] = res.sweeps;
    return 0;
}

#include "../../american_monte_carlo_b200/csrc/philox.cuh"
extern "C" void amc_test_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                uint32_t* out) {
    amc::Philox4 r = amc::philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

#include "../../american_monte_carlo_b200/csrc/gbm_quad.cuh"
extern "C" int amc_test_fixed_point_bits(double drift_log2, double vol_log2, int n_steps) {
    return amc::fixed_point_bits(drift_log2, vol_log2, n_steps);
}
