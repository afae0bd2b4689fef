"""GPU tier: the path-free ("lean") sweep against the stored float path set of the same seed.

The float generator keeps each path's log2-price as an exact int32 fixed-point sum; the lean sweep walks it backwards
(L_{t-1} = L_t - q_t, q_t regenerated from the Philox counter), so it must see THE SAME prices as the stored matrix:
identical columns, identical exercise step for every path, price equal to summation-order rounding.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MKT = dict(S0=36.0, r=0.06, sigma=0.2, T=1.0)


def both(amc, n, P, seed):
    stored = amc.generate_asset_paths(MKT["S0"], MKT["r"], MKT["sigma"], MKT["T"], n, P, rng="philox", seed=seed,
                                      dtype="float32")
    lean = amc.generate_asset_paths(MKT["S0"], MKT["r"], MKT["sigma"], MKT["T"], n, P, rng="philox", seed=seed,
                                    dtype="float32", store_paths=False)
    return stored, lean


def test_lean_set_regenerates_the_stored_columns_bit_for_bit(amc):
    n, P = 37, 10_003                                           # ragged: the last quad is partial
    stored, lean = both(amc, n, P, 5)
    assert lean.nbytes_device <= 4 * (P + 32) and stored.nbytes_device >= 4 * P * (n + 1)
    for t in (0, 1, 2, n // 2, n - 1, n):
        np.testing.assert_array_equal(lean.column(t), stored.column(t))
    np.testing.assert_array_equal(lean.rows(17, 230), stored.rows(17, 230))
    np.testing.assert_array_equal(np.asarray(lean[P - 5:]), np.asarray(stored[P - 5:]))
    np.testing.assert_array_equal(amc.precompute_barrier_hit_matrix(lean, 33.0), amc.precompute_barrier_hit_matrix(stored, 33.0))


@pytest.mark.parametrize("kw", [dict(), dict(state_dtype="float32"), dict(barrier_level=33.0), dict(exercise_type="European"),
                                dict(basis_type="Chebyshev", degree=5, scaling=True), dict(option_type="Call", K=34.0)])
def test_lean_sweep_takes_the_stored_sweep_decisions(amc, kw):
    n, P = 50, 200_003
    stored, lean = both(amc, n, P, 11)
    args = dict(K=40.0, option_type="Put", barrier_level=None, exercise_type="American", basis_type="Power", degree=3)
    extra = {}
    for k, v in kw.items():
        (args if k in args else extra)[k] = v
    call = (args["K"], MKT["r"], MKT["T"] / n, args["option_type"], args["barrier_level"], args["exercise_type"],
            args["basis_type"], args["degree"])
    a = amc.lsm_price(stored, *call, **extra, want_exercise_steps=True, want_regression=True)
    b = amc.lsm_price(lean, *call, **extra, want_exercise_steps=True, want_regression=True)
    assert int((a.exercise_steps != b.exercise_steps).sum()) == 0
    assert abs(a.price - b.price) <= 1e-11 * max(abs(a.price), 1e-12)
    np.testing.assert_allclose(b.gamma, a.gamma, rtol=1e-8, atol=1e-10)
    # the lean set is not consumed by a sweep: pricing it again gives the same number
    assert amc.lsm_price(lean, *call, **extra).price == b.price


def test_lean_drop_in_entry_points(amc):
    n, P = 20, 50_000
    stored, lean = both(amc, n, P, 3)
    pa, ca = amc.lsmc_option_pricing(stored, 40.0, MKT["r"], MKT["T"] / n, "Put", None, "American", "Power", 3)
    pb, cb = amc.lsmc_option_pricing(lean, 40.0, MKT["r"], MKT["T"] / n, "Put", None, "American", "Power", 3)
    assert abs(pa - pb) <= 1e-11 * pa
    for t in (0, 7, n):
        np.testing.assert_array_equal(cb[t][1], ca[t][1])
        np.testing.assert_allclose(cb[t][2], ca[t][2], rtol=1e-8, atol=1e-9)
    ea, eb = amc.compute_ccr_exposures(ca), amc.compute_ccr_exposures(cb)
    np.testing.assert_allclose(np.array(eb)[:, 1:], np.array(ea)[:, 1:], rtol=1e-8, atol=1e-9)
    with pytest.raises(ValueError):
        amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, n, P, rng="philox", seed=3, dtype="float64", store_paths=False)
    with pytest.raises(ValueError):
        amc.lsm_price_batch(lean, [(40.0, "Put", "American"), (41.0, "Put", "American")], MKT["r"], MKT["T"] / n)


def test_lean_prices_a_set_larger_than_the_stored_mode_could_hold_per_byte(amc):
    """4M paths x 252 steps: the stored float matrix is 4 GB, the lean state 16 MB + the sweep's own state."""
    n, P = 252, 4_000_000
    lean = amc.generate_asset_paths(MKT["S0"], MKT["r"], MKT["sigma"], MKT["T"], n, P, rng="philox", seed=42,
                                    dtype="float32", store_paths=False)
    assert lean.nbytes_device < 17_000_000
    res = amc.lsm_price(lean, 40.0, MKT["r"], MKT["T"] / n, "Put", None, "American", "Power", 3, want_cashflows=True,
                        state_dtype="float32")
    se = res.cashflow0.std() / np.sqrt(P)
    assert abs(res.price - 4.4847) < 5 * se + 0.004           # reference at 1M x 252 (another sample): 4.48475
