"""Per-phase clock trace of the one-cluster sweep kernel (AMC_CLUSTER_TRACE=1): one configuration per process,
`python scripts/r2x_cluster_trace.py PATHS BASIS DEGREE [DTYPE]`; libamc prints the mean SM cycles of each phase of a pass
on the 8th pricing."""
import os
import sys

os.environ["AMC_CLUSTER_TRACE"] = "1"
os.environ.setdefault("AMC_CLUSTER_MAX_PATHS", "1000000000")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import american_monte_carlo_b200 as amc  # noqa: E402

P, basis, deg = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
dtype = sys.argv[4] if len(sys.argv) > 4 else "float64"
n = 50
dp = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, n, P, rng="philox", seed=7, dtype=dtype)
for _ in range(10):
    r = amc.lsm_price(dp, 40.0, 0.06, 1.0 / n, "Put", None, "American", basis, deg,
                      state_dtype="float32" if dtype == "float32" else "float64")
print(P, basis, deg, dtype, "kind", r.timing["sweep_kind"], "total_ms", r.timing["total_ms"], flush=True)
