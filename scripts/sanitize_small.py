"""Every kernel of libamc on small shapes, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import american_monte_carlo_b200 as amc  # noqa: E402
from american_monte_carlo_b200 import sweeps  # noqa: E402


def main():
    S0, K, r, sigma, T = 36.0, 40.0, 0.06, 0.2, 1.0
    for n, P in ((7, 1537), (12, 5001), (3, 33)):
        rng = np.random.default_rng(n)
        Z = rng.standard_normal((P, n))
        for dtype in ("float64", "float32"):
            dp = amc.paths_from_normals(Z, S0, r, sigma, T, dtype=dtype)
            for kw in (dict(), dict(scaling=True)):
                for barrier in (None, 33.0):
                    for basis, deg in (("Power", 3), ("Chebyshev", 4), ("Legendre", 8), ("Power", 0)):
                        a = amc.lsm_price(dp, K, r, T / n, "Put", barrier, "American", basis, deg, want_exercise_steps=True,
                                          want_cashflows=True, want_svd=True, **kw)
                        assert np.isfinite(a.price)
            if dtype == "float32":
                amc.lsm_price(dp, K, r, T / n, "Put", None, "American", "Power", 3, state_dtype="float32", want_cashflows=True)
            b = amc.lsm_price_batch(dp, [(40.0, "Put", "American"), (38.0, "Call", "European"), (44.0, "Put", "American")],
                                    r, T / n, None, "Power", 3, want_gamma=True,
                                    state_dtype="float32" if dtype == "float32" else "float64")
            assert np.all(np.isfinite(b[0]))
            price, cont = amc.lsmc_option_pricing(dp, K, r, T / n, "Put", None, "European", "Chebyshev", 4)
            _ = cont[1], cont[n], np.asarray(dp), dp[0], dp[:, 1]
            ex = amc.compute_ccr_exposures(cont)
            assert len(ex) == n + 1
            amc.precompute_barrier_hit_matrix(dp, 33.0)
            dp.free()
        for dtype in ("float64", "float32"):
            dq = amc.generate_asset_paths(S0, r, sigma, T, n, P, rng="philox", seed=3, dtype=dtype)
            amc.lsm_price(dq, K, r, T / n, "Call", None, "American", "Laguerre", 5, scaling=True)
            host = np.asarray(dq)
            dq.free()
            dh = amc.paths_from_host(host, dtype=dtype)
            amc.lsm_price(dh, K, r, T / n, "Put", None, "American", "Power", 2)
            dh.free()
        # path-free set: the cooperative sweep kernel, every mode, against the stored set of the same seed
        ds = amc.generate_asset_paths(S0, r, sigma, T, n, P, rng="philox", seed=5, dtype="float32")
        dl = amc.generate_asset_paths(S0, r, sigma, T, n, P, rng="philox", seed=5, dtype="float32", store_paths=False)
        for kw in (dict(), dict(state_dtype="float32"), dict(scaling=True)):
            for barrier in (None, 33.0):
                for ex_type in ("American", "European"):
                    for basis, deg in (("Power", 3), ("Legendre", 8), ("Power", 0)):
                        a = amc.lsm_price(ds, K, r, T / n, "Put", barrier, ex_type, basis, deg, want_exercise_steps=True,
                                          want_regression=True, **kw)
                        b = amc.lsm_price(dl, K, r, T / n, "Put", barrier, ex_type, basis, deg, want_exercise_steps=True,
                                          want_regression=True, **kw)
                        assert np.isfinite(b.price) and abs(a.price - b.price) <= 1e-10 * max(abs(a.price), 1e-12)
                        assert (a.exercise_steps == b.exercise_steps).all()
        assert np.array_equal(np.asarray(dl), np.asarray(ds))
        _, cont = amc.lsmc_option_pricing(dl, K, r, T / n, "Put", None, "American", "Chebyshev", 4)
        assert len(amc.compute_ccr_exposures(cont)) == n + 1
        ds.free()
        dl.free()
    x = np.linspace(30, 50, 1001)
    amc.intrinsic_value(x, 40.0, "Put")
    amc.get_basis_polynomials(x, "Legendre", 5)
    amc.regression_estimate(x, np.maximum(40 - x, 0), "Power", 3)
    amc.compute_ccr_exposures([(0, None, np.random.default_rng(1).standard_normal(3001))])
    g = sweeps.contract_grid(36.0, 0.06, [38.0, 42.0], [0.2, 0.3], [0.5], 6, 3000, seed=1)
    assert g.shape == (2, 2, 1)
    print("sanitize_small: ok")


if __name__ == "__main__":
    main()
