"""GPU tier: the expectations of the reference's own test file against the drop-in module.

/root/reference/unit_test.py holds two tests: twelve LSMC-vs-QuantLib comparisons ({Put, Call} x {European, American} x
barrier at {none, 80 %, 60 %} of spot; spot = strike = 100, T = 1, r = 1 %, sigma = 20 %, 100 steps, 10 000 paths, seed 42,
Chebyshev degree 4; both prices rounded to 4 decimals, tolerance 0.2 -- unit_test.py:6-25,29-50) and a payoff known-answer
test (unit_test.py:54-62).  They are re-created here through the shim `american_monte_carlo` at the repository root (CUDA hot
path + QuantLib-free benchmark stand-in); /root/reference does not exist on the GPU box.  On top of the reference's
tolerance each LSMC price must equal the number the reference itself produces (tests/golden/golden.json, "ut_*") to 1e-10.
One case fails IN THE REFERENCE: American call without barrier, |8.1894 - 8.4135| = 0.224 > 0.2 (rank-truncated Chebyshev-4
fit); a faithful drop-in reproduces the price and therefore the failure, so it is a strict xfail.
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MARKET = dict(S0=100, K=100, T=1.0, r=0.01, sigma=0.2)
GRID = dict(n_time_steps=100, n_paths=10000)
REFERENCE_FAILS = {("Call", "American", None)}


def _cases():
    for barrier_pct, exercise, side in itertools.product((None, 80, 60), ("European", "American"), ("Put", "Call")):
        key = (side, exercise, barrier_pct)
        marks = [pytest.mark.xfail(strict=True, reason="the reference's own price misses its own tolerance here")] \
            if key in REFERENCE_FAILS else []
        yield pytest.param(*key, marks=marks, id=f"{exercise}-{side}-barrier{barrier_pct}")


@pytest.mark.parametrize("side, exercise, barrier_pct", list(_cases()))
def test_lsmc_price_against_benchmark_and_reference_number(amc, golden, side, exercise, barrier_pct):
    import american_monte_carlo as dropin
    barrier = MARKET["S0"] * barrier_pct / 100 if barrier_pct else None
    n, P = GRID["n_time_steps"], GRID["n_paths"]
    np.random.seed(42)
    paths = dropin.generate_asset_paths(MARKET["S0"], MARKET["r"], MARKET["sigma"], MARKET["T"], n, P)
    price, _ = dropin.lsmc_option_pricing(paths, MARKET["K"], MARKET["r"], MARKET["T"] / n, side, barrier, exercise,
                                          "Chebyshev", 4)
    bench = dropin.get_quantlib_option(MARKET["S0"], MARKET["K"], MARKET["r"], MARKET["T"], MARKET["sigma"], n, side,
                                       exercise, barrier).NPV()
    want = golden[f"ut_{side}_{exercise}_{barrier_pct}"]["price"]
    assert abs(float(price) - want) <= 1e-10 * max(abs(want), 1e-12) + 1e-14
    assert abs(round(float(price), 4) - round(bench, 4)) < 0.2


def test_payoff_known_answers(amc):
    import american_monte_carlo as dropin
    spots = np.array([90, 100, 110])
    np.testing.assert_array_almost_equal(dropin.intrinsic_value(spots, 100, "Put"), [10, 0, 0])
    np.testing.assert_array_almost_equal(dropin.intrinsic_value(spots, 100, "Call"), [0, 0, 10])
