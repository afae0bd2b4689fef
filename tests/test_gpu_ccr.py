"""GPU tier: exposures on the device (compute_ccr_exposures, amc.py:400-414) -- radix-select percentiles + mean."""
import json
import os

import numpy as np
import pytest

from oracle import lsm_oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    with open(os.path.join(HERE, "golden", "ccr_golden.json")) as f:
        return json.load(f)["cases"]


def test_percentile_select_is_exact_on_host_arrays(amc):
    """amc_percentiles against np.percentile / np.mean on arbitrary arrays: the order statistics are exact and the
    interpolation follows numpy's _lerp, so the percentiles are bit-identical; the mean differs only by summation order."""
    rng = np.random.default_rng(5)
    arrays = [rng.standard_normal(100_003), np.abs(rng.standard_normal(4097)) * 1e-300, rng.integers(0, 5, 999).astype(float),
              np.array([3.5]), np.array([2.0, -1.0]), np.concatenate([rng.standard_normal(1000), [np.nan, np.inf, -np.inf]]),
              -np.abs(rng.standard_normal(50_000)) * 1e6, np.zeros(777), np.array([-0.0, 0.0, 1e-320, -1e-320])]
    for a in arrays:
        got = amc.compute_ccr_exposures([(0, None, a)])[0]
        ok = a[np.isfinite(a)]
        assert got[1] == np.percentile(ok, 5) and got[2] == np.percentile(ok, 95), (len(a), got)
        assert abs(got[3] - ok.mean()) <= 1e-13 * max(np.abs(ok).max(), 1e-300)
    empty = amc.compute_ccr_exposures([(3, None, np.array([np.nan, np.inf]))])[0]
    assert empty[0] == 3 and all(np.isnan(v) for v in empty[1:])                    # amc.py:405-408


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_device_exposures_match_reference_golden(amc, idx):
    """Lazy continuation values -> exposures entirely on the device, against the tuples the reference produced.
    The continuation values themselves differ from numpy's fitted values by ~cond(A)*eps, so the
    percentiles agree to that tolerance; against percentiles of the device's own materialised vectors they are exact."""
    c = _cases()[idx]
    dt = c["T"] / c["n_time_steps"]
    np.random.seed(c["seed"])
    Z = orc.draw_normals(c["n_paths"], c["n_time_steps"])
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"])
    price, cont = amc.lsmc_option_pricing(dp, c["K"], c["r"], dt, c["option_type"], c["barrier_level"],
                                          c["exercise_type"], c["basis_type"], c["degree"], **c["kwargs"])
    assert abs(price - c["price"]) <= 1e-10 * max(abs(c["price"]), 1e-12)
    got = amc.compute_ccr_exposures(cont)
    assert len(got) == c["n_time_steps"] + 1
    scale = max(abs(w[2]) for w in c["exposures"]) + 1e-12
    for (t, a, b, m), want in zip(got, c["exposures"]):
        assert t == want[0]
        assert abs(a - want[1]) <= 2e-7 * scale and abs(b - want[2]) <= 2e-7 * scale and abs(m - want[3]) <= 2e-7 * scale
    # exact against the vectors the same object materialises
    for t in (0, 1, c["n_time_steps"] // 2, c["n_time_steps"]):
        v = cont[t][2]
        assert got[t][1] == np.percentile(v, 5) and got[t][2] == np.percentile(v, 95)
        assert abs(got[t][3] - v.mean()) <= 1e-13 * max(np.abs(v).max(), 1e-300)
    assert got[-1][1:] == (0.0, 0.0, 0.0)                                            # zeros at maturity, amc.py:145


def test_device_exposures_float32_paths(amc):
    dp = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 10, 200_001, rng="philox", seed=9, dtype="float32")
    price, cont = amc.lsmc_option_pricing(dp, 40.0, 0.06, 0.1, "Put", None, "American", "Power", 3)
    got = amc.compute_ccr_exposures(cont)
    for t in (1, 5, 9):
        v = cont[t][2]
        assert got[t][1] == np.percentile(v, 5) and got[t][2] == np.percentile(v, 95)
