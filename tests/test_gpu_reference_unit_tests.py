"""GPU tier: the reference's own test file (/root/reference/unit_test.py) re-stated against the drop-in module.

Same imports, same parametrisation, same assertion (unit_test.py:3,6-25,29-50,54-62); `american_monte_carlo` here is
the shim at the repository root (CUDA hot path + QuantLib-free benchmark stand-in).  /root/reference does not exist on
the GPU box, hence a restatement rather than running the file itself.  One case is a known failure OF THE REFERENCE:
Call / American / no barrier gives |8.1894 - 8.4135| = 0.224 > 0.2 with the reference's own LSMC price (rank-truncated
Chebyshev-4 fit); a faithful drop-in reproduces the price and therefore the failure.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# full-precision LSMC prices produced by the reference itself, seed 42 (tests/golden/golden.json, "ut_*")
REFERENCE_LSMC = {("Put", "European", None): "ut_Put_European_None", ("Call", "European", None): "ut_Call_European_None",
                  ("Put", "American", None): "ut_Put_American_None", ("Call", "American", None): "ut_Call_American_None",
                  ("Put", "European", 80): "ut_Put_European_80", ("Call", "European", 80): "ut_Call_European_80",
                  ("Put", "American", 80): "ut_Put_American_80", ("Call", "American", 80): "ut_Call_American_80",
                  ("Put", "European", 60): "ut_Put_European_60", ("Call", "European", 60): "ut_Call_European_60",
                  ("Put", "American", 60): "ut_Put_American_60", ("Call", "American", 60): "ut_Call_American_60"}


def run_lsmc_quantlib_test(S0, K, T, r, sigma, n_time_steps, n_paths, option_type, exercise_type, barrier_level):
    from american_monte_carlo import generate_asset_paths, get_quantlib_option, lsmc_option_pricing
    np.random.seed(42)
    dt = T / n_time_steps
    basis_type, degree = "Chebyshev", 4
    paths = generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths)
    lsmc_price, _ = lsmc_option_pricing(paths, K, r, dt, option_type, barrier_level, exercise_type, basis_type, degree)
    full = float(lsmc_price)
    lsmc_price = round(lsmc_price, 4)
    quantlib_option = get_quantlib_option(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type, barrier_level)
    quantlib_price = round(quantlib_option.NPV(), 4)
    return full, lsmc_price, quantlib_price


@pytest.mark.parametrize("option_type, exercise_type, barrier_pct", [
    ("Put", "European", None),
    ("Call", "European", None),
    ("Put", "American", None),
    pytest.param("Call", "American", None, marks=pytest.mark.xfail(
        strict=True, reason="fails in the reference too: its own LSMC price 8.1894 is 0.224 from the benchmark 8.4135")),
    ("Put", "European", 80),
    ("Call", "European", 80),
    ("Put", "American", 80),
    ("Call", "American", 80),
    ("Put", "European", 60),
    ("Call", "European", 60),
    ("Put", "American", 60),
    ("Call", "American", 60),
])
def test_lsmc_quantlib_comparison(amc, golden, option_type, exercise_type, barrier_pct):
    S0, K, T, r, sigma = 100, 100, 1.0, 0.01, 0.2
    n_time_steps, n_paths = 100, 10000
    barrier_level = S0 * barrier_pct / 100 if barrier_pct else None
    full, lsmc_price, quantlib_price = run_lsmc_quantlib_test(S0, K, T, r, sigma, n_time_steps, n_paths, option_type,
                                                              exercise_type, barrier_level)
    want = golden[REFERENCE_LSMC[(option_type, exercise_type, barrier_pct)]]["price"]
    assert abs(full - want) <= 1e-10 * max(abs(want), 1e-12) + 1e-14          # the reference's own number
    assert abs(lsmc_price - quantlib_price) < 0.2                              # unit_test.py:21


def test_intrinsic_value(amc):
    from american_monte_carlo import intrinsic_value
    S = np.array([90, 100, 110])
    K = 100
    np.testing.assert_array_almost_equal(intrinsic_value(S, K, "Put"), [10, 0, 0])
    np.testing.assert_array_almost_equal(intrinsic_value(S, K, "Call"), [0, 0, 10])
