"""CPU tier: the QuantLib-free benchmark stand-in against the three QuantLib values stored in the reference notebook
(AmericanMonteCarlo.ipynb:203, 249, 313) and against the 12 unit_test.py cases' LSMC goldens at the test's own
tolerance (unit_test.py:21: |round(lsmc,4) - round(benchmark,4)| < 0.2)."""
import pytest

from american_monte_carlo_b200.benchmarks import get_quantlib_option

NB = dict(S0=95, K=100, r=0.01, T=1.0, sigma=0.2)


def test_notebook_quantlib_values():
    assert f"{get_quantlib_option(**NB, n_steps=100, option_type='Put', exercise_type='European').NPV():.4f}" == "9.8928"
    assert f"{get_quantlib_option(**NB, n_steps=100, option_type='Put', exercise_type='American').NPV():.4f}" == "10.0198"
    assert f"{get_quantlib_option(**NB, n_steps=100, option_type='Put', exercise_type='European', barrier_level=70).NPV():.4f}" == "4.0316"


def test_error_behaviour_matches_reference():
    with pytest.raises(NotImplementedError):
        get_quantlib_option(**NB, exercise_type="Bermudan", barrier_level=70)          # amc.py:45
    with pytest.raises(KeyError):
        get_quantlib_option(**NB, exercise_type="Bermudan")                            # amc.py:53
    with pytest.raises(RuntimeError):
        get_quantlib_option(60, 100, 0.01, 1.0, 0.2, barrier_level=70).NPV()           # touched: amc.py:219 relies on it


def test_unit_test_cases_against_lsmc_goldens(golden):
    # the reference's own test most likely fails Call/American/None (0.224 > 0.2) because of the
    # rank-truncated Chebyshev-4 fit; a faithful drop-in reproduces that, it is recorded, not "fixed".
    known_reference_failures = {"ut_Call_American_None"}
    for name, c in golden.items():
        if not name.startswith("ut_"):
            continue
        bench = get_quantlib_option(c["S0"], c["K"], c["r"], c["T"], c["sigma"], c["n_time_steps"], c["option_type"],
                                    c["exercise_type"], c["barrier_level"]).NPV()
        diff = abs(round(c["price"], 4) - round(bench, 4))
        if name in known_reference_failures:
            assert 0.2 <= diff < 0.3, (name, diff)
        else:
            assert diff < 0.2, (name, c["price"], bench)
