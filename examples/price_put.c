/* Minimal C host of libamc (include/amc.h): price the Longstaff-Schwartz Table-1 American put on device-generated
 * Philox paths, then the same strike ladder as one batch.  Plain C, no CUDA headers needed on the host side.
 *
 *   gcc -O2 -I../include price_put.c -L../american_monte_carlo_b200 -lamc -Wl,-rpath,'$ORIGIN/../american_monte_carlo_b200' -o price_put
 *   (libamc.so is linked as  -l:libamc.so  when it keeps its in-tree name; see examples/Makefile)
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "amc.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != AMC_OK) {                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, amc_last_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void) {
    const double S0 = 36.0, K = 40.0, r = 0.06, sigma = 0.2, T = 1.0;
    const int n = 50;
    const int64_t P = 1000000;

    amc_ctx* ctx = NULL;
    CHECK(amc_ctx_create(0, NULL, &ctx));                     /* fails without a CUDA device: there is no CPU path */

    amc_paths* paths = NULL;                                  /* replaces generate_asset_paths (amc.py:72-81) */
    CHECK(amc_paths_generate(ctx, S0, r, sigma, T, n, P, 0, P, AMC_F32, 42u, &paths));

    amc_lsm_spec spec;                                        /* replaces the arguments of lsmc_option_pricing (amc.py:180) */
    memset(&spec, 0, sizeof(spec));
    spec.K = K; spec.r = r; spec.dt = T / n; spec.barrier = NAN; spec.scaling_factor = 2.0;
    spec.is_put = 1; spec.is_american = 1; spec.basis = AMC_BASIS_POWER; spec.degree = 3;
    spec.state_f32 = 1;

    double price = 0.0;
    amc_lsm_timing tm;
    CHECK(amc_lsm_price(ctx, paths, &spec, &price, NULL, NULL, NULL, &tm, 0));
    printf("American put K=%.0f: %.5f  (sweep %.3f ms, %d + %d launches)\n", K, price, tm.total_ms, tm.step_launches,
           tm.solve_launches);

    enum { NK = 5 };
    amc_lsm_spec ladder[NK];
    double prices[NK];
    for (int i = 0; i < NK; ++i) { ladder[i] = spec; ladder[i].K = 36.0 + 2.0 * i; }
    CHECK(amc_lsm_price_batch(ctx, paths, ladder, NK, prices, NULL, NULL, 0));
    for (int i = 0; i < NK; ++i) printf("  K=%.0f  %.5f\n", ladder[i].K, prices[i]);

    /* the same contract on a path-free set: no path matrix (4 bytes per path instead of 4 (n+1)), same price */
    amc_paths* lean = NULL;
    double lean_price = 0.0;
    CHECK(amc_paths_generate_lean(ctx, S0, r, sigma, T, n, P, 0, P, 42u, &lean));
    CHECK(amc_lsm_price(ctx, lean, &spec, &lean_price, NULL, NULL, NULL, NULL, 0));
    printf("path-free set: %.5f (difference %.1e)\n", lean_price, lean_price - price);
    CHECK(amc_paths_free(lean));

    CHECK(amc_paths_free(paths));
    CHECK(amc_ctx_destroy(ctx));
    return 0;
}
