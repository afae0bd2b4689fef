"""GPU tier: the sweep drivers (numeric part of the reference's american_monte_carlo_additional_plots.py) and the
batched strike x vol x maturity grid of BASELINE.json config 4."""
import numpy as np
import pytest

from oracle import lsm_oracle as orc

pytestmark = pytest.mark.gpu


def test_contract_grid_equals_per_contract_pricing(amc):
    from american_monte_carlo_b200 import sweeps
    strikes, vols, mats = [36.0, 40.0, 44.0], [0.2, 0.4], [0.5, 1.0]
    n, P = 20, 50_000
    grid = sweeps.contract_grid(36.0, 0.06, strikes, vols, mats, n, P, seed=5, dtype="float64")
    assert grid.shape == (3, 2, 2)
    cell = 0
    for iv, sg in enumerate(vols):
        for im, T in enumerate(mats):
            dp = sweeps._generate_unsharded(amc.default_context(), 36.0, 0.06, sg, T, n, P, "float64", 5 + cell)
            host_paths = np.asarray(dp)
            for ik, K in enumerate(strikes):
                solo = amc.lsm_price(dp, K, 0.06, T / n, "Put", None, "American", "Power", 3).price
                assert abs(grid[ik, iv, im] - solo) <= 1e-12 * max(solo, 1.0)
            # and against the CPU oracle on the very same (device-generated) paths, one strike per cell
            want = orc.lsm_backward(host_paths, 40.0, 0.06, T / n, "Put", None, "American", "Power", 3,
                                    keep_continuation=False).price
            assert abs(grid[1, iv, im] - want) <= 1e-10 * want
            dp.free()
            cell += 1
    # put prices increase with the strike and with volatility
    assert np.all(np.diff(grid, axis=0) > 0) and np.all(np.diff(grid, axis=1) > 0)


def test_plot_sweeps_numeric_part(amc):
    from american_monte_carlo_b200 import sweeps
    np.random.seed(42)
    common = dict(S0=100, K=100, r=0.05, T=1.0, sigma=0.2, option_type="Put", exercise_type="American", barrier_level=None)
    ns, prices, bench = sweeps.convergence_with_paths(n_time_steps=50, path_range=[2000, 20000], **common)
    assert ns == [2000, 20000] and prices.shape == (2,) and abs(prices[1] - bench) < 0.25
    ts, prices_t, bench_t = sweeps.convergence_with_time_steps(n_paths=20000, time_step_range=[10, 50], **common)
    assert prices_t.shape == (2,) and abs(prices_t[1] - bench_t) < 0.25
    err, best = sweeps.error_heatmap(time_step_range=[10, 25], path_range=[2000, 10000], **common)
    assert err.shape == (2, 2) and best[0] in (2000, 10000) and best[1] in (10, 25)
    # one path set, every basis and degree: the oracle on the same paths must agree (unscaled bases truncate ranks)
    np.random.seed(3)
    degs, by_basis, _ = sweeps.error_vs_basis_degree(n_time_steps=20, n_paths=5000, max_degree=5, **common)
    np.random.seed(3)
    paths = orc.generate_asset_paths(100, 0.05, 0.2, 1.0, 20, 5000)
    for basis, got in by_basis.items():
        for d in degs:
            want = orc.lsm_backward(paths, 100, 0.05, 1.0 / 20, "Put", None, "American", basis, d,
                                    keep_continuation=False).price
            assert abs(got[d] - want) <= 1e-9 * want, (basis, d, got[d], want)
