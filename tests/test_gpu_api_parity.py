"""GPU tier: the remaining names of the reference's module surface -- apply_exercise,
estimate_continuation_values, perform_backward_iteration, main -- against the oracle / reference semantics."""
import numpy as np
import pytest

from oracle import lsm_oracle as orc

pytestmark = pytest.mark.gpu


def test_apply_exercise_in_place(amc):
    rng = np.random.default_rng(1)
    P, m = 5000, 1700
    cf, tau = rng.random(P), np.full(P, 50, dtype=np.int64)
    idx = np.sort(rng.choice(P, m, replace=False))
    ev, ce = rng.random(m), rng.random(m)
    ce[:5] = ev[:5]                                        # ties must NOT exercise (strict >, amc.py:91)
    want_cf, want_tau = cf.copy(), tau.copy()
    mask = ev > ce
    want_cf[idx[mask]] = ev[mask]
    want_tau[idx[mask]] = 7
    amc.apply_exercise(cf, tau, idx, ev, ce, 7)
    np.testing.assert_array_equal(cf, want_cf)
    np.testing.assert_array_equal(tau, want_tau)
    with pytest.raises(IndexError):
        amc.apply_exercise(cf, tau, np.array([P]), ev[:1], ce[:1], 7)
    amc.apply_exercise(cf, tau, idx[:0], ev[:0], ce[:0], 7)       # empty candidate set (amc.py:158 guards it)


@pytest.mark.parametrize("basis,degree,kwargs", [("Power", 3, {}), ("Chebyshev", 4, dict(scaling=True, scaling_factor=1))])
def test_estimate_continuation_values(amc, basis, degree, kwargs):
    np.random.seed(11)
    paths = orc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 10, 20000)
    n = 10
    cf = np.maximum(40.0 - paths[:, n], 0)
    tau = np.full(len(cf), n)
    want = orc.continuation_estimate(paths, 6, 0.06, 0.1, cf, tau, basis, degree, **kwargs)
    got = amc.estimate_continuation_values(paths, 6, 0.06, 0.1, cf, tau, basis, degree, **kwargs)
    assert got.min() >= 0.0
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-9 * want.max())
    with pytest.raises(ValueError):
        amc.estimate_continuation_values(paths, 6, 0.06, 0.1, cf, tau, "Hermite", 3)


@pytest.mark.parametrize("opt,ex,barrier", [("Put", "American", None), ("Put", "American", 33.0), ("Call", "European", None)])
def test_perform_backward_iteration_contract(amc, opt, ex, barrier):
    """Same in-place contract as amc.py:139-167: cashflows, exercise_times, the (reversed) list of per-step tuples."""
    np.random.seed(21)
    n, P = 12, 8000
    paths = orc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, n, P)
    want = orc.lsm_backward(paths, 40.0, 0.06, 1.0 / n, opt, barrier, ex, "Power", 3)
    cashflows, exercise_times, cont = np.zeros(P), np.full(P, n), []
    hit = amc.precompute_barrier_hit_matrix(paths, barrier)
    np.testing.assert_array_equal(hit, orc.knock_in_flags(paths, barrier))
    amc.perform_backward_iteration(40.0, 0.06, 1.0 / n, n, hit, cashflows, paths, opt, exercise_times, ex, cont, "Power", 3)
    np.testing.assert_array_equal(exercise_times, want.exercise_times)
    np.testing.assert_array_equal(cashflows, want.cashflows)              # payoff at the exercise step: exact
    price = np.mean(cashflows * np.exp(-0.06 * (1.0 / n) * exercise_times))   # amc.py:196
    assert abs(price - want.price) <= 1e-12 * max(want.price, 1e-12)
    assert [c[0] for c in cont] == list(range(n + 1))
    for (t, S_t, cv), (tw, Sw, cw) in zip(cont, want.continuation_values):
        np.testing.assert_array_equal(S_t, Sw)
        np.testing.assert_allclose(cv, cw, rtol=0, atol=5e-9 * max(cw.max(), 1.0))
    bad = hit.copy()
    bad[0, :] = False
    bad[0, 3] = True                                                       # not a running OR
    with pytest.raises(NotImplementedError):
        amc.perform_backward_iteration(40.0, 0.06, 1.0 / n, n, bad, cashflows, paths, opt, exercise_times, ex, [], "Power", 3)


def test_main_prints_reference_lines(amc, capsys, golden):
    """The notebook's American put (AmericanMonteCarlo.ipynb:248-249): LSMC 10.3838, QuantLib 10.0198."""
    np.random.seed(42)
    params = dict(S0=95, K=100, T=1.0, r=0.01, sigma=0.2, n_time_steps=100, n_paths=1000, option_type="Put",
                  exercise_type="American", barrier_level=None, basis_type="Chebyshev", degree=10, scaling=True,
                  scaling_factor=1, n_plotted_paths=100, difference_type="difference", vmin_diff=None, vmax_diff=None)
    out = amc.main(params)
    printed = capsys.readouterr().out.splitlines()
    assert printed[0] == "American Put Option Price without Barrier (LSMC): 10.3838"
    assert printed[1] == "American Put Option Price without Barrier (QuantLib): 10.0198"
    assert len(out["lsmc_ccr_exposures"]) == 101 and out["lsmc_ccr_exposures"][-1][1:] == (0.0, 0.0, 0.0)
    np.random.seed(42)
    params.update(exercise_type="European", barrier_level=70)
    amc.main(params)
    printed = capsys.readouterr().out.splitlines()
    assert printed[0] == "European Put Option Price with Barrier at 70 (LSMC): 4.0108"          # ipynb:312-313
    assert printed[1] == "European Put Option Price with Barrier at 70 (QuantLib): 4.0316"
    assert printed[2] == "European Put Option Price without Barrier (QuantLib): 9.8928"


def test_plain_c_host_prices_through_the_abi(libamc_path, tmp_path):
    """examples/price_put.c: a C program with no Python in the loop -- context, Philox paths, one contract, a batch."""
    import os
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(libamc_path)
    exe = tmp_path / "price_put"
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "price_put.c"), "-L", libdir,
                    "-l:libamc.so", "-lm", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    price = float(lines[0].split(":")[1].split()[0])
    assert abs(price - 4.472) < 0.03                                   # Longstaff-Schwartz Table 1
    ladder = [float(l.split()[1]) for l in lines[1:6]]
    assert len(ladder) == 5 and all(b > a for a, b in zip(ladder, ladder[1:]))     # put price increases with the strike
    assert abs(ladder[2] - price) <= 1e-9 * price                      # K = 40 inside the batch == the single contract
    lean = float(lines[6].split(":")[1].split()[0])                    # path-free set of the same seed: the same price
    assert lines[6].startswith("path-free set") and abs(lean - price) <= 2e-5 * price   # printed to 5 decimals
