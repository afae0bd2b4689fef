"""Generate tests/golden/ccr_golden.json from the UNMODIFIED reference (build container only).

    python tests/golden/make_ccr_golden.py

Runs the reference's own generate_asset_paths + lsmc_option_pricing + compute_ccr_exposures
(/root/reference/american_monte_carlo.py:400-414) on small seeded cases, refuses to write unless
oracle/lsm_oracle.py::ccr_exposures is bit-identical, and stores the (t, PFE_5, PFE_95, EPE) tuples.
/root/reference does not exist on the GPU box; tests only read the JSON.
"""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for _m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.gridspec", "QuantLib"]:
    sys.modules[_m] = MagicMock()
sys.path.insert(0, "/root/reference")
import american_monte_carlo as ref  # noqa: E402  (the real reference)

from oracle import lsm_oracle as orc  # noqa: E402

CASES = [
    dict(name="ccr_put_american_power3", S0=36.0, K=40.0, r=0.06, sigma=0.2, T=1.0, n_time_steps=20, n_paths=20000,
         option_type="Put", exercise_type="American", barrier_level=None, basis_type="Power", degree=3, kwargs={}, seed=42),
    dict(name="ccr_put_european_cheb4_scaled", S0=95, K=100, r=0.01, sigma=0.2, T=1.0, n_time_steps=25, n_paths=5001,
         option_type="Put", exercise_type="European", barrier_level=None, basis_type="Chebyshev", degree=4,
         kwargs=dict(scaling=True, scaling_factor=1), seed=42),
    dict(name="ccr_call_american_di80_legendre3", S0=100, K=100, r=0.01, sigma=0.2, T=1.0, n_time_steps=16, n_paths=10000,
         option_type="Call", exercise_type="American", barrier_level=80, basis_type="Legendre", degree=3,
         kwargs=dict(scaling=True), seed=7),
]


def main():
    out = []
    for c in CASES:
        dt = c["T"] / c["n_time_steps"]
        np.random.seed(c["seed"])
        paths = ref.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], c["n_time_steps"], c["n_paths"])
        price, cont = ref.lsmc_option_pricing(paths, c["K"], c["r"], dt, c["option_type"], c["barrier_level"],
                                              c["exercise_type"], c["basis_type"], c["degree"], **c["kwargs"])
        want = ref.compute_ccr_exposures(cont)
        np.random.seed(c["seed"])
        opaths = orc.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], c["n_time_steps"], c["n_paths"])
        oprice, ocont = orc.lsmc_option_pricing(opaths, c["K"], c["r"], dt, c["option_type"], c["barrier_level"],
                                                c["exercise_type"], c["basis_type"], c["degree"], **c["kwargs"])
        got = orc.ccr_exposures(ocont)
        assert oprice == price
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert a[0] == b[0] and all(x == y or (np.isnan(x) and np.isnan(y)) for x, y in zip(a[1:], b[1:])), (a, b)
        out.append(dict(c, price=float(price), exposures=[[int(t), float(a), float(b), float(m)] for t, a, b, m in want]))
        print(c["name"], "ok", want[1])
    with open(os.path.join(HERE, "ccr_golden.json"), "w") as f:
        json.dump(dict(source="/root/reference/american_monte_carlo.py:400-414 (compute_ccr_exposures)",
                       numpy=np.__version__, cases=out), f, indent=1)


if __name__ == "__main__":
    main()
