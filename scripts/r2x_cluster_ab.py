"""A/B of the one-cluster sweep kernel against the launch chain on small path sets (run once per setting of AMC_CLUSTER;
the switch is read once per process).  Prints one JSON line per (paths, storage, basis/degree): the median and the minimum
of the device time of a whole backward sweep (CUDA events inside amc_lsm_price) over `reps` pricings of the same set."""
import json
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import american_monte_carlo_b200 as amc  # noqa: E402


def main():
    reps = int(os.environ.get("AB_REPS", "40"))
    n = 50
    cases = [("float64", "float64", "Power", 3, {}), ("float32", "float32", "Power", 3, {}),
             ("float64", "float64", "Chebyshev", 4, {}), ("float64", "float64", "Laguerre", 8, dict(scaling=True))]
    for P in (10_000, 25_000, 50_000, 75_000, 100_000, 140_000):
        for dtype, state, basis, deg, kw in cases:
            dp = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, n, P, rng="philox", seed=7, dtype=dtype)
            ms, kinds, price = [], set(), None
            for i in range(reps + 5):
                r = amc.lsm_price(dp, 40.0, 0.06, 1.0 / n, "Put", None, "American", basis, deg, state_dtype=state, **kw)
                if i >= 5:
                    ms.append(r.timing["total_ms"])
                kinds.add(r.timing["sweep_kind"])
                price = float(r.price)
            dp.free()
            print(json.dumps(dict(paths=P, steps=n, dtype=dtype, state=state, basis=basis, degree=deg, kinds=sorted(kinds),
                                  median_ms=statistics.median(ms), min_ms=min(ms), us_per_step=1e3 * statistics.median(ms) / (n + 1),
                                  price=price, cluster_env=os.environ.get("AMC_CLUSTER", "default"))), flush=True)


if __name__ == "__main__":
    main()
