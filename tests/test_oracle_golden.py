"""CPU tier: the oracle against the committed golden vectors (which were produced by the unmodified reference)."""
import os

import numpy as np
import pytest

from conftest import oracle_case
from oracle import lsm_oracle as orc

SMALL = 100_000


def small_cases(golden):
    return [c for c in golden.values() if c["n_paths"] <= SMALL]


def test_golden_has_reference_known_answers(golden):
    # prices printed in the reference notebook (AmericanMonteCarlo.ipynb:202,248,312,377,442)
    printed = {"nb_european_put": "9.9485", "nb_american_put": "10.3838", "nb_di70_european_put": "4.0108",
               "nb_di70_european_put_200x10000": "3.7835", "nb_di70_european_put_unscaled": "4.0108"}
    for name, txt in printed.items():
        assert f"{golden[name]['price']:.4f}" == txt
    assert golden["c1_power3"]["price"] == 4.4783398987704475
    assert golden["c2_power3_10M"]["price"] == 4.475181386178888      # BASELINE.md, lstsq rank 3 at t=1
    assert golden["c2_power3_10M"]["ranks"][1] == 3
    # the reduced shapes of configs[2] and configs[4], from the reference itself
    assert f"{golden['c3_reduced']['price']:.10f}" == "4.4847469992"
    assert f"{golden['c5_reduced']['price']:.10f}" == "4.4890595342"


def test_big_exercise_step_fixture_matches_its_goldens(golden):
    steps = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big_exercise_steps.npz"))
    for name in ("c3_reduced", "c5_reduced", "c3_reduced_f32paths", "c5_reduced_f32paths"):
        c = golden[name]
        tau = steps[name]
        assert tau.shape == (c["n_paths"],) and tau.dtype == np.uint8
        assert np.bincount(tau, minlength=c["n_time_steps"] + 1).tolist() == c["exercise_step_hist"]


def test_oracle_reproduces_every_small_golden_case_bitwise(golden):
    for c in small_cases(golden):
        _, paths, res = oracle_case(c)
        n = c["n_time_steps"]
        assert res.price == c["price"], c["name"]
        assert float(np.sum(res.cashflows)) == c["cashflow_sum"], c["name"]
        assert np.bincount(res.exercise_times, minlength=n + 1).tolist() == c["exercise_step_hist"], c["name"]
        assert [res.steps[t]["rank"] for t in range(n)] == c["ranks"], c["name"]
        assert [float(paths[:, n].sum()), float(paths[:, n // 2].sum()), float(paths.min()), float(paths.max())] \
            == c["path_checksum"], c["name"]
        for t, rec in c.get("steps", {}).items():
            np.testing.assert_array_equal(res.steps[int(t)]["sv"], rec["sv"])
            np.testing.assert_array_equal(res.steps[int(t)]["coeffs"], rec["coeffs"])
            np.testing.assert_array_equal(res.continuation_values[int(t)][2][:8], c["cont_probe"][t])


def test_intrinsic_value_known_answer():
    # unit_test.py:54-62
    S = np.array([90, 100, 110])
    np.testing.assert_array_almost_equal(orc.payoff(S, 100, "Put"), [10, 0, 0])
    np.testing.assert_array_almost_equal(orc.payoff(S, 100, "Call"), [0, 0, 10])
    np.testing.assert_array_almost_equal(orc.payoff(S, 100, "anything else"), [0, 0, 10])   # amc.py:86


def test_unknown_basis_raises_value_error():
    with pytest.raises(ValueError, match="Unknown basis type"):
        orc.basis_matrix(np.array([1.0, 2.0]), "Hermite", 2)


def test_exercise_type_other_than_american_never_exercises(golden):
    c = golden["small_degree1_scaled"]
    np.random.seed(1)
    paths = orc.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], 8, 500)
    a = orc.lsm_backward(paths, c["K"], c["r"], c["T"] / 8, "Put", None, "Bermudan", "Power", 2)
    b = orc.lsm_backward(paths, c["K"], c["r"], c["T"] / 8, "Put", None, "European", "Power", 2)
    assert a.price == b.price and (a.exercise_times == 8).all()


def test_oracle_matches_reference_when_reference_is_present():
    """Build container only: re-pin the oracle against the real reference (skipped on the GPU box)."""
    if not os.path.exists("/root/reference/american_monte_carlo.py"):
        pytest.skip("reference not mounted")
    import sys
    from unittest.mock import MagicMock
    saved = dict(sys.modules)
    try:
        for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.gridspec",
                  "QuantLib"]:
            sys.modules[m] = MagicMock()
        sys.modules.pop("american_monte_carlo", None)
        sys.path.insert(0, "/root/reference")
        import american_monte_carlo as ref
        assert ref.__file__.startswith("/root/reference")
        for seed, (ot, et, b, basis, d, kw) in enumerate([("Put", "American", None, "Chebyshev", 4, {}),
                                                           ("Call", "American", 90.0, "Legendre", 5, dict(scaling=True)),
                                                           ("Put", "European", 85.0, "Power", 2, {})]):
            np.random.seed(seed)
            p_ref = ref.generate_asset_paths(100, 0.03, 0.25, 1.0, 30, 4000)
            np.random.seed(seed)
            p_orc = orc.generate_asset_paths(100, 0.03, 0.25, 1.0, 30, 4000)
            assert np.array_equal(p_ref, p_orc)
            pr, cr = ref.lsmc_option_pricing(p_ref, 100, 0.03, 1 / 30, ot, b, et, basis, d, **kw)
            po, co = orc.lsmc_option_pricing(p_orc, 100, 0.03, 1 / 30, ot, b, et, basis, d, **kw)
            assert pr == po
            for (ta, sa, ca), (tb, sb, cb) in zip(cr, co):
                assert ta == tb and np.array_equal(sa, sb) and np.array_equal(ca, cb)
    finally:
        sys.path.remove("/root/reference")
        sys.modules.pop("american_monte_carlo", None)
        for m in list(sys.modules):
            if m not in saved:
                sys.modules.pop(m, None)


def test_oracle_ccr_exposures_match_reference_golden():
    """compute_ccr_exposures (amc.py:400-414): the oracle against tuples produced by the reference itself
    (tests/golden/make_ccr_golden.py)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ccr_golden.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 3
    for c in cases:
        dt = c["T"] / c["n_time_steps"]
        np.random.seed(c["seed"])
        paths = orc.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], c["n_time_steps"], c["n_paths"])
        price, cont = orc.lsmc_option_pricing(paths, c["K"], c["r"], dt, c["option_type"], c["barrier_level"],
                                              c["exercise_type"], c["basis_type"], c["degree"], **c["kwargs"])
        assert price == c["price"]
        got = orc.ccr_exposures(cont)
        assert len(got) == c["n_time_steps"] + 1
        for (t, a, b, m), want in zip(got, c["exposures"]):
            assert [t, a, b, m] == want
