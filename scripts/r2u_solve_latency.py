"""Sweep time per step for the reference's default regression (Chebyshev-4, unscaled, S0 = 100: lstsq rank 4 of 5 on half
the steps -> the SVD path of the solve) against certified configurations, at small path counts (solve-dominated)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import american_monte_carlo_b200 as amc

for P in (10_000, 100_000):
    for basis, deg, kw, S0 in (("Chebyshev", 4, {}, 100.0), ("Chebyshev", 4, dict(scaling=True), 100.0), ("Power", 3, {}, 36.0),
                               ("Power", 5, {}, 36.0), ("Power", 5, dict(scaling=True), 36.0), ("Legendre", 8, {}, 36.0),
                               ("Legendre", 8, dict(scaling=True), 36.0), ("Chebyshev", 10, {}, 100.0)):
        n = 100
        dp = amc.generate_asset_paths(S0, 0.01, 0.2, 1.0, n, P, rng="philox", seed=1)
        best = None
        for rep in range(6):
            res = amc.lsm_price(dp, S0, 0.01, 1.0 / n, "Put", None, "American", basis, deg, **kw)
            t = res.timing["total_ms"]
            best = t if best is None else min(best, t)
        cert = int((res.rank[:n] == deg + 1).sum())
        print(f"P={P:7d} {basis:9s} d={deg:2d} scaling={bool(kw)!s:5s} sweep {best:7.3f} ms = {1e3 * best / (n + 1):6.1f} us/step; "
              f"full-rank steps {cert}/{n}, price {res.price:.4f}")
        dp.free()
