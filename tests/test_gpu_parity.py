"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the golden vectors.

Tolerances (BASELINE.json north_star): identical injected normals -> price within 1e-10 relative in FP64 and NOT ONE
path exercising at a different step; 1e-5 relative with FP32 path storage; Monte Carlo standard error with the
independent Philox stream.
"""
import numpy as np
import pytest

from conftest import oracle_case
from oracle import lsm_oracle as orc

pytestmark = pytest.mark.gpu

SMALL = ["nb_european_put", "nb_american_put", "nb_di70_european_put", "nb_di70_european_put_200x10000",
         "nb_di70_european_put_unscaled", "ut_Put_European_None", "ut_Call_European_None", "ut_Put_American_None",
         "ut_Call_American_None", "ut_Put_European_80", "ut_Call_European_80", "ut_Put_American_80",
         "ut_Call_American_80", "ut_Put_European_60", "ut_Call_European_60", "ut_Put_American_60",
         "ut_Call_American_60", "c1_power3", "c1_chebyshev3", "c1_legendre3", "c1_call_power3", "c1_di30_power3",
         "small_power8_unscaled", "small_legendre8_scaled", "small_chebyshev10_unscaled", "small_degree0",
         "small_degree1_scaled", "deep_itm_exercise_at_0", "tiny_paths_lt_k"]


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def price_args(c):
    dt = c["T"] / c["n_time_steps"]
    return (c["K"], c["r"], dt, c["option_type"], c["barrier_level"], c["exercise_type"], c["basis_type"], c["degree"])


@pytest.mark.parametrize("name", SMALL)
def test_injected_normals_price_decisions_and_ranks(amc, golden, name):
    """K1z + sweep on the reference's own normals: paths, price, every exercise decision, every lstsq rank."""
    c = golden[name]
    Z, paths, o = oracle_case(c)
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"])
    assert dp.shape == paths.shape
    got_paths = np.asarray(dp)
    assert np.max(np.abs(got_paths - paths) / paths) < 1e-13         # log-space sum vs cumprod: ~3e-15
    res = amc.lsm_price(dp, *price_args(c), **c["kwargs"], want_exercise_steps=True, want_cashflows=True,
                        want_regression=True, want_svd=True)
    n = c["n_time_steps"]
    flips = int((res.exercise_steps != o.exercise_times).sum())
    assert flips == 0, f"{flips} paths exercise at a different step"
    assert rel(res.price, c["price"]) <= 1e-10 or abs(res.price - c["price"]) <= 1e-14
    assert res.rank[:n].tolist() == c["ranks"]
    np.testing.assert_allclose(res.cashflow0, o.cashflows * np.exp(-c["r"] * (c["T"] / n) * o.exercise_times),
                               rtol=1e-12, atol=1e-12 * c["K"])       # K - S cancels: paths agree to ~3e-15 * S
    # singular values of the design matrix, where numpy kept them
    for t, rec in c.get("steps", {}).items():
        t = int(t)
        r = rec["rank"]
        if res.sv[t, 1] > 0 or r == 1:          # degenerate columns (fewer distinct points than k) report s_1 only
            np.testing.assert_allclose(res.sv[t, :r], rec["sv"][:r], rtol=1e-6)
    dp.free()


@pytest.mark.parametrize("name", ["ut_Put_American_80", "c1_power3", "small_power8_unscaled", "nb_american_put",
                                  "tiny_paths_lt_k"])
def test_adopted_reference_paths(amc, golden, name):
    """amc_paths_from_host: the oracle's own [P, n+1] matrix is adopted (measured column maps) -> bit-equal paths."""
    c = golden[name]
    _, paths, o = oracle_case(c)
    dp = amc.paths_from_host(paths)
    np.testing.assert_array_equal(np.asarray(dp), paths)
    np.testing.assert_array_equal(dp[:, c["n_time_steps"] // 2], paths[:, c["n_time_steps"] // 2])
    np.testing.assert_array_equal(dp[1], paths[1])
    mu, sg = dp.column_maps()
    np.testing.assert_allclose(mu, paths.mean(axis=0), rtol=1e-12)
    res = amc.lsm_price(dp, *price_args(c), **c["kwargs"], want_exercise_steps=True)
    assert int((res.exercise_steps != o.exercise_times).sum()) == 0
    assert rel(res.price, c["price"]) <= 1e-10
    # the reference-shaped call on a plain ndarray
    p2, _ = amc.lsmc_option_pricing(paths, *price_args(c), **c["kwargs"])
    assert rel(p2, c["price"]) <= 1e-10


def test_drop_in_call_sequence_reproduces_notebook_and_unit_test_prices(amc, golden):
    """np.random.seed + generate_asset_paths + lsmc_option_pricing exactly as unit_test.py:7-12 / the notebook."""
    for name in ["nb_american_put", "nb_european_put", "nb_di70_european_put_200x10000", "ut_Put_American_None",
                 "ut_Call_American_60"]:
        c = golden[name]
        np.random.seed(c["seed"])
        paths = amc.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], c["n_time_steps"], c["n_paths"])
        price, cont = amc.lsmc_option_pricing(paths, *price_args(c), **c["kwargs"])
        assert rel(price, c["price"]) <= 1e-10 or abs(price - c["price"]) < 1e-14
        assert len(cont) == c["n_time_steps"] + 1
        if "printed_in_notebook" in c:
            assert f"{price:.4f}" == c["printed_in_notebook"]


@pytest.mark.parametrize("name", ["nb_american_put", "ut_Put_European_80", "c1_power3", "small_legendre8_scaled"])
def test_lazy_continuation_values(amc, golden, name):
    c = golden[name]
    Z, paths, o = oracle_case(c)
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"])
    _, cont = amc.lsmc_option_pricing(dp, *price_args(c), **c["kwargs"])
    n = c["n_time_steps"]
    for t in sorted({0, 1, n // 2, n - 1, n}):
        tt, S_t, cv = cont[t]
        t_o, S_o, cv_o = o.continuation_values[t]
        assert tt == t_o == t
        np.testing.assert_allclose(S_t, S_o, rtol=1e-13)
        scale = max(np.abs(cv_o).max(), 1e-12)
        # numpy's own fitted values carry an error ~ cond(kept part of A) * eps (LAPACK gelsd); the CUDA solver was
        # checked against the exact truncated projection to 1e-14 (tests/test_solver_host.py, DESIGN.md "Solver").
        tol = 1e-9
        if t < n:
            d = o.steps[t]
            tol = max(tol, 20 * d["sv"][0] / d["sv"][d["rank"] - 1] * 2.2e-16)
        assert np.abs(cv - cv_o).max() <= tol * scale, (t, tol)
    assert (cont[-1][2] == 0).all()


def test_fp32_path_storage_within_1e5(amc, golden):
    c = golden["c1_power3"]
    Z, paths, o = oracle_case(c)
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"], dtype="float32")
    got = np.asarray(dp)
    assert np.max(np.abs(got - paths) / paths) < 1e-7
    np.testing.assert_array_equal(got, paths.astype(np.float32).astype(np.float64))   # f64 arithmetic, one rounding
    res = amc.lsm_price(dp, *price_args(c), want_exercise_steps=True)
    assert rel(res.price, c["price"]) <= 1e-5
    assert (res.exercise_steps != o.exercise_times).mean() < 1e-3


@pytest.mark.parametrize("name", ["c1_power3", "ut_Put_American_80", "small_legendre8_scaled", "c1_call_power3"])
def test_fp32_state_within_1e5(amc, golden, name):
    """state_f32: the per-path state (cashflow discounted to time 0) is stored as float; sums, solve and the exercise
    test stay in double.  Against the FP64 oracle the FP32 tolerance (1e-5) applies; against the same sweep with a
    double state the difference is the state's rounding only."""
    c = golden[name]
    Z, paths, o = oracle_case(c)
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"], dtype="float32")
    r64 = amc.lsm_price(dp, *price_args(c), **c["kwargs"], want_exercise_steps=True, want_cashflows=True)
    r32 = amc.lsm_price(dp, *price_args(c), **c["kwargs"], want_exercise_steps=True, want_cashflows=True,
                        state_dtype="float32")
    assert rel(r32.price, c["price"]) <= 1e-5
    assert rel(r32.price, r64.price) <= 2e-7
    assert (r32.exercise_steps != r64.exercise_steps).mean() < 1e-4
    same = r32.exercise_steps == r64.exercise_steps
    np.testing.assert_allclose(r32.cashflow0[same], r64.cashflow0[same], rtol=1e-6, atol=0)
    # batched sweep with a float state
    b32 = amc.lsm_price_batch(dp, [(c["K"], c["option_type"], c["exercise_type"])] * 3, c["r"],
                              c["T"] / c["n_time_steps"], c["barrier_level"], c["basis_type"], c["degree"],
                              state_dtype="float32", **c["kwargs"])
    assert np.all(np.abs(b32 - r32.price) <= 1e-9 * max(abs(r32.price), 1e-3))
    dp.free()
    d64 = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"])
    with pytest.raises(ValueError):
        amc.lsm_price(d64, *price_args(c), **c["kwargs"], state_dtype="float32")
    d64.free()


def test_philox_paths_statistics_and_price_within_mc_error(amc, golden):
    c = golden["c1_power3"]
    P, n = 400_000, c["n_time_steps"]
    for dtype in ("float64", "float32"):
        dp = amc.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], n, P, rng="philox", seed=1234, dtype=dtype)
        ST = dp.column(n)
        logret = np.log(ST / c["S0"])
        m = (c["r"] - 0.5 * c["sigma"] ** 2) * c["T"]
        assert abs(logret.mean() - m) < 5 * c["sigma"] / np.sqrt(P)
        assert abs(logret.std() - c["sigma"]) < 5 * c["sigma"] / np.sqrt(2 * P)
        # one-step increments are i.i.d. normal: check skewness / kurtosis of a middle step
        inc = np.log(dp.column(n // 2 + 1) / dp.column(n // 2))
        zs = (inc - inc.mean()) / inc.std()
        assert abs((zs ** 3).mean()) < 5 * np.sqrt(6 / P) and abs((zs ** 4).mean() - 3) < 5 * np.sqrt(24 / P)
        assert (dp.column(0) == c["S0"]).all()
        res = amc.lsm_price(dp, *price_args(c), want_cashflows=True)
        se = res.cashflow0.std() / np.sqrt(P)
        # the golden (100k reference paths) carries its own Monte Carlo error
        assert abs(res.price - c["price"]) < 4 * se * np.sqrt(1 + P / c["n_paths"])
        assert abs(res.price - 4.472) < 0.03                                    # Longstaff-Schwartz Table 1: 4.472
        # determinism and seed sensitivity
        dp2 = amc.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], n, P, rng="philox", seed=1234, dtype=dtype)
        assert amc.lsm_price(dp2, *price_args(c)).price == res.price
        dp3 = amc.generate_asset_paths(c["S0"], c["r"], c["sigma"], c["T"], n, P, rng="philox", seed=1235, dtype=dtype)
        assert amc.lsm_price(dp3, *price_args(c)).price != res.price


def test_philox_f32_log_sum_accuracy(amc):
    """The float path generator keeps the cumulative log2-price as an exact int32 fixed-point sum (gbm_quad.cuh): with
    sigma = 0 the path is deterministic, S_t = S0 exp(r t), and must be met to a few float roundings at every one of 252
    steps -- the error is the price formation's (int -> float, ex2.approx, one multiply: ~2.5 ulp), it does not grow
    along the path; with sigma > 0 the mean of S_t must match the forward within Monte Carlo error at the last step."""
    n = 252
    dp = amc.generate_asset_paths(36.0, 0.06, 0.0, 1.0, n, 4096, rng="philox", seed=3, dtype="float32")
    A = np.asarray(dp)
    want = 36.0 * np.exp(0.06 * np.arange(n + 1) / n)
    err = np.abs(A - want[None, :]) / want[None, :]
    assert err.max() < 4e-7
    dp.free()
    P = 2_000_000
    dq = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, n, P, rng="philox", seed=3, dtype="float32")
    ST = dq.column(n)
    fwd = 36.0 * np.exp(0.06)
    assert abs(ST.mean() - fwd) < 5 * ST.std() / np.sqrt(P)
    lr = np.log(ST / 36.0)
    assert abs(lr.mean() - (0.06 - 0.02)) < 5 * 0.2 / np.sqrt(P) and abs(lr.std() - 0.2) < 5 * 0.2 / np.sqrt(2 * P)
    dq.free()


def test_philox_paths_do_not_depend_on_sharding(amc):
    """Counters are GLOBAL path ids: a shard generated with an offset equals the same rows of the full set."""
    import ctypes as C
    from american_monte_carlo_b200 import _native as N
    ctx = amc.default_context()
    full = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 13, 1003, rng="philox", seed=7)
    A = np.asarray(full)
    for dtype in (N.F64, N.F32):
        ref = A if dtype == N.F64 else np.asarray(amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 13, 1003,
                                                                          rng="philox", seed=7, dtype="float32"))
        h = C.c_void_p()
        N.check(N.lib().amc_paths_generate(ctx.handle, 36.0, 0.06, 0.2, 1.0, 13, 400, 301, 1003, dtype,
                                           C.c_uint64(7), C.byref(h)))
        shard = amc.DevicePaths(ctx, h, 400, 1003, 13, dtype, 301)
        np.testing.assert_array_equal(np.asarray(shard), ref[301:701])


def test_small_array_ops(amc):
    # unit_test.py:54-62
    S = np.array([90, 100, 110])
    np.testing.assert_array_almost_equal(amc.intrinsic_value(S, 100, "Put"), [10, 0, 0])
    np.testing.assert_array_almost_equal(amc.intrinsic_value(S, 100, "Call"), [0, 0, 10])
    rng = np.random.default_rng(0)
    X = 100 * np.exp(0.2 * rng.standard_normal(30000))
    Y = np.maximum(100 - X, 0) + rng.standard_normal(X.size)
    for basis, deg, kw in [("Power", 3, {}), ("Chebyshev", 4, {}), ("Legendre", 6, dict(scaling=True)),
                           ("Chebyshev", 10, dict(scaling=True, scaling_factor=1))]:
        diag = {}
        want = orc.regression_fit(X, Y, basis, deg, diag=diag, **kw)
        got = amc.regression_estimate(X, Y, basis, deg, **kw)
        tol = max(1e-9, 20 * diag["sv"][0] / diag["sv"][diag["rank"] - 1] * 2.2e-16)   # numpy's own error
        assert np.abs(got - want).max() <= tol * np.abs(want).max()
        U_ = (X - diag["centre"]) / (kw.get("scaling_factor", 2) * diag["spread"]) if kw.get("scaling") else X
        Us = np.linalg.svd(orc.basis_matrix(U_, basis, deg), full_matrices=False)[0][:, :diag["rank"]]
        exact = Us @ (Us.T @ Y)                                                       # exact projection
        assert np.abs(got - exact).max() <= max(1e-9, tol / 10) * np.abs(want).max()
        A = amc.get_basis_polynomials(X[:100] / 100, basis, deg)
        np.testing.assert_allclose(A, orc.basis_matrix(X[:100] / 100, basis, deg), rtol=1e-11, atol=1e-13)
    paths = orc.generate_asset_paths(100, 0.01, 0.2, 1.0, 20, 500)
    for b in (None, 90.0, 10.0, 1e9):
        np.testing.assert_array_equal(amc.precompute_barrier_hit_matrix(paths, b), orc.knock_in_flags(paths, b))


def test_edge_cases(amc):
    # no early exercise for unknown exercise types (amc.py:154), call for unknown option types (amc.py:86)
    np.random.seed(3)
    paths = orc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 10, 2000)
    dp = amc.paths_from_host(paths)
    a = amc.lsm_price(dp, 40.0, 0.06, 0.1, "Put", None, "Bermudan", "Power", 3).price
    b = amc.lsm_price(dp, 40.0, 0.06, 0.1, "Put", None, "European", "Power", 3).price
    assert a == b and rel(a, orc.lsm_backward(paths, 40.0, 0.06, 0.1, "Put", None, "European", "Power", 3).price) < 1e-13
    a = amc.lsm_price(dp, 40.0, 0.06, 0.1, "Straddle", None, "American", "Power", 3).price
    assert rel(a, orc.lsm_backward(paths, 40.0, 0.06, 0.1, "Call", None, "American", "Power", 3).price) < 1e-10
    # barrier never hit -> 0; barrier always hit -> vanilla
    assert amc.lsm_price(dp, 40.0, 0.06, 0.1, "Put", 1.0, "American", "Power", 3).price == 0.0
    v = amc.lsm_price(dp, 40.0, 0.06, 0.1, "Put", None, "American", "Power", 3).price
    assert amc.lsm_price(dp, 40.0, 0.06, 0.1, "Put", 1e6, "American", "Power", 3).price == v
    # all out of the money -> 0
    assert amc.lsm_price(dp, 1.0, 0.06, 0.1, "Put", None, "American", "Power", 3).price == 0.0
    # sigma = 0: every column is constant, every regression is the rank-1 mean fit
    flat = orc.paths_from_normals(np.zeros((50, 6)), 36.0, 0.06, 0.0, 1.0)
    want = orc.lsm_backward(flat, 40.0, 0.06, 1 / 6, "Put", None, "American", "Power", 3).price
    assert rel(amc.lsmc_option_pricing(flat, 40.0, 0.06, 1 / 6, "Put", None, "American", "Power", 3)[0], want) < 1e-12
    # odd path counts and a single path
    for P in (1, 2, 3, 255, 257):
        np.random.seed(P)
        pp = orc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 5, P)
        want = orc.lsm_backward(pp, 40.0, 0.06, 0.2, "Put", None, "American", "Power", 2).price
        got = amc.lsmc_option_pricing(pp, 40.0, 0.06, 0.2, "Put", None, "American", "Power", 2)[0]
        assert rel(got, want) < 1e-9 or abs(got - want) < 1e-12, (P, got, want)
    with pytest.raises(ValueError, match="Unknown basis type"):
        amc.lsmc_option_pricing(dp, 40.0, 0.06, 0.1, "Put", None, "American", "Hermite", 3)
    with pytest.raises(ValueError):
        amc.lsmc_option_pricing(dp, 40.0, 0.06, 0.1, "Put", None, "American", "Power", 11)


def test_laguerre_extension_matches_oracle_extension(amc, golden):
    c = golden["small_legendre8_scaled"]
    Z, paths, _ = oracle_case(c)
    dt = c["T"] / c["n_time_steps"]
    want = orc.lsm_backward(paths, c["K"], c["r"], dt, "Put", None, "American", "Laguerre", 8, scaling=True)
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"])
    res = amc.lsm_price(dp, c["K"], c["r"], dt, "Put", None, "American", "Laguerre", 8, scaling=True,
                        want_exercise_steps=True)
    assert int((res.exercise_steps != want.exercise_times).sum()) == 0
    assert rel(res.price, want.price) <= 1e-10


@pytest.mark.slow
def test_config2_10M_paths_fp64_injected_normals(amc, golden):
    """BASELINE.json configs[1]: 10M x 50, FP64, the reference's own seed-42 normals (drawn here on the host, 4 GB).
    Size-independent checks: price to 1e-10 of the golden, the histogram of exercise steps identical (a flipped
    decision moves a path between bins), numpy's rank-3 truncation at t=1 reproduced."""
    c = golden["c2_power3_10M"]
    np.random.seed(c["seed"])
    Z = orc.draw_normals(c["n_paths"], c["n_time_steps"])
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"])
    del Z
    n = c["n_time_steps"]
    col = dp.column(n)
    assert rel(col.sum(), c["path_checksum"][0]) < 1e-13
    res = amc.lsm_price(dp, *price_args(c), want_exercise_steps=True, want_cashflows=True)
    assert res.rank[:n].tolist() == c["ranks"] and res.rank[1] == 3
    hist = np.bincount(res.exercise_steps, minlength=n + 1).tolist()
    moved = sum(abs(a - b) for a, b in zip(hist, c["exercise_step_hist"]))
    assert moved == 0, f"exercise-step histogram differs by {moved}"
    assert int(np.count_nonzero(res.cashflow0)) == c["n_nonzero_cashflows"]
    assert rel(res.price, c["price"]) <= 1e-10


# ------------------------------------------------------------------------------------------------ contract batches
def _batch_contracts():
    # strike axis of a sweep, both payoff sides, both exercise styles -- all on one path set
    return [(40.0, "Put", "American"), (36.0, "Put", "American"), (44.0, "Put", "American"), (38.0, "Call", "American"),
            (40.0, "Put", "European"), (34.0, "Call", "European"), (32.0, "Put", "American")]


@pytest.mark.parametrize("basis,degree,kwargs", [("Power", 3, {}), ("Chebyshev", 4, {}),
                                                 ("Legendre", 8, dict(scaling=True, scaling_factor=2))])
def test_batch_equals_oracle_per_contract(amc, basis, degree, kwargs):
    """amc_lsm_price_batch: every contract of a batch must reproduce the oracle's price for that contract alone
    (identical injected normals, 1e-10) and the price amc_lsm_price gives for it (rounding only)."""
    S0, r, sigma, T, n, P = 36.0, 0.06, 0.2, 1.0, 40, 60_000
    np.random.seed(123)
    Z = orc.draw_normals(P, n)
    paths = orc.paths_from_normals(Z, S0, r, sigma, T)
    dp = amc.paths_from_normals(Z, S0, r, sigma, T)
    contracts = _batch_contracts()
    got, gamma = amc.lsm_price_batch(dp, contracts, r, T / n, None, basis, degree, want_gamma=True, **kwargs)
    assert got.shape == (len(contracts),) and gamma.shape == (len(contracts), n + 1, 11)
    for i, (K, opt, ex) in enumerate(contracts):
        want = orc.lsm_backward(paths, K, r, T / n, opt, None, ex, basis, degree, keep_continuation=False, **kwargs)
        solo = amc.lsm_price(dp, K, r, T / n, opt, None, ex, basis, degree, **kwargs)
        assert rel(got[i], want.price) <= 1e-10 or abs(got[i] - want.price) <= 1e-14, (i, got[i], want.price)
        assert rel(got[i], solo.price) <= 1e-12 or abs(got[i] - solo.price) <= 1e-14
        if ex == "American":
            np.testing.assert_allclose(gamma[i, 1:n], solo.gamma[1:n], rtol=1e-7, atol=1e-9)
    dp.free()


def test_batch_with_barrier_and_f32_paths(amc):
    S0, r, sigma, T, n, P = 100.0, 0.01, 0.2, 1.0, 30, 50_000
    np.random.seed(7)
    Z = orc.draw_normals(P, n)
    paths = orc.paths_from_normals(Z, S0, r, sigma, T)
    dp = amc.paths_from_normals(Z, S0, r, sigma, T)
    contracts = [(100.0, "Put", "American"), (95.0, "Put", "American"), (105.0, "Put", "European")]
    got = amc.lsm_price_batch(dp, contracts, r, T / n, 80.0, "Power", 3, scaling=True)
    for i, (K, opt, ex) in enumerate(contracts):
        want = orc.lsm_backward(paths, K, r, T / n, opt, 80.0, ex, "Power", 3, keep_continuation=False, scaling=True)
        assert rel(got[i], want.price) <= 1e-10
    dp.free()
    d32 = amc.paths_from_normals(Z, S0, r, sigma, T, dtype="float32")
    got32 = amc.lsm_price_batch(d32, contracts, r, T / n, None, "Power", 3, scaling=True)
    for i, (K, opt, ex) in enumerate(contracts):
        want = orc.lsm_backward(paths, K, r, T / n, opt, None, ex, "Power", 3, keep_continuation=False, scaling=True)
        assert rel(got32[i], want.price) <= 1e-5                       # FP32 path storage tolerance (north_star)
    d32.free()


def test_batch_argument_errors(amc):
    dp = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 5, 1000, rng="philox", seed=1)
    with pytest.raises(ValueError):
        amc.lsm_price_batch(dp, [(40.0, "Put", "American")] * 300, 0.06, 0.2, None, "Power", 3)     # > AMC_MAX_BATCH
    with pytest.raises(ValueError):
        amc.lsm_price_batch(dp, [(40.0, "Put", "American")], 0.06, 0.2, None, "Hermite", 3)
    assert amc.lsm_price_batch(dp, [], 0.06, 0.2).shape == (0,)
    one = amc.lsm_price_batch(dp, [(40.0, "Put", "American")], 0.06, 0.2, None, "Power", 3)
    solo = amc.lsm_price(dp, 40.0, 0.06, 0.2, "Put", None, "American", "Power", 3)
    assert one[0] == solo.price
    dp.free()


def test_repeated_sweeps_replay_a_graph_with_identical_results(amc):
    """The launch chain of a sweep is captured into a CUDA graph on its second identical occurrence and replayed
    afterwards (api.cu); every occurrence must give the same bits, and a different contract in between must not be
    served by the stale graph."""
    dp = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 20, 30_000, rng="philox", seed=5)
    args = (40.0, 0.06, 0.05, "Put", None, "American", "Power", 3)
    prices = [amc.lsm_price(dp, *args).price for _ in range(5)]
    assert len(set(prices)) == 1
    other = amc.lsm_price(dp, 42.0, *args[1:]).price
    assert other != prices[0]
    again = [amc.lsm_price(dp, *args).price for _ in range(3)]
    assert set(again) == {prices[0]}
    batch = [amc.lsm_price_batch(dp, [(40.0, "Put", "American"), (42.0, "Put", "American")], 0.06, 0.05, None, "Power", 3)
             for _ in range(4)]
    for b in batch[1:]:
        np.testing.assert_array_equal(b, batch[0])
    assert abs(batch[0][0] - prices[0]) <= 1e-12 * prices[0] and abs(batch[0][1] - other) <= 1e-12 * other
    dp.free()


def test_host_arrays_are_streamed_in_chunks_with_identical_results(amc, monkeypatch):
    """Host normals / adopted host matrices go through two staging halves in chunks (copy stream || kernel): many small
    chunks must give bit-identical path matrices and prices to a single chunk."""
    rng = np.random.default_rng(17)
    P, n = 70_001, 20
    Z = rng.standard_normal((P, n))
    monkeypatch.delenv("AMC_STAGE_CHUNK_MB", raising=False)
    one = amc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0)
    A = np.asarray(one)
    price = amc.lsm_price(one, 40.0, 0.06, 0.05, "Put", None, "American", "Power", 3).price
    monkeypatch.setenv("AMC_STAGE_CHUNK_MB", "1")                     # 1 MB chunks: ~11 chunks for Z, ~12 for the matrix
    many = amc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0)
    np.testing.assert_array_equal(np.asarray(many), A)
    assert amc.lsm_price(many, 40.0, 0.06, 0.05, "Put", None, "American", "Power", 3).price == price
    adopted = amc.paths_from_host(A)
    np.testing.assert_array_equal(np.asarray(adopted), A)
    adopted32 = amc.paths_from_host(A, dtype="float32")
    np.testing.assert_array_equal(np.asarray(adopted32), A.astype(np.float32).astype(np.float64))
    for d in (one, many, adopted, adopted32):
        d.free()


@pytest.mark.parametrize("name", ["small_power8_unscaled", "small_chebyshev10_unscaled", "small_legendre8_scaled",
                                  "ut_Put_American_None", "ut_Call_American_80", "nb_american_put"])
def test_rank_truncated_steps_through_the_production_solve(amc, golden, name):
    """Without want_svd the solve takes its production route: the full-rank certificate, and -- from degree 6 -- the
    warp-parallel Jacobi SVD for the steps that numpy truncates (degree-8 / degree-10 unscaled bases: every step).  Ranks
    per step, every exercise decision and the price must be the reference's, and the singular values reported for the
    truncated steps must be numpy's."""
    c = golden[name]
    Z, paths, o = oracle_case(c)
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"])
    res = amc.lsm_price(dp, *price_args(c), **c["kwargs"], want_exercise_steps=True, want_regression=True)
    dp.free()
    n = c["n_time_steps"]
    assert int((res.exercise_steps != o.exercise_times).sum()) == 0
    assert res.rank[:n].tolist() == c["ranks"]
    assert rel(res.price, c["price"]) <= 1e-10
    for t, rec in c.get("steps", {}).items():
        t = int(t)
        r = rec["rank"]
        if r < c["degree"] + 1 and c["degree"] >= 6 and res.sv[t, 1] > 0:      # truncated step solved by the warp routine
            np.testing.assert_allclose(res.sv[t, :r], rec["sv"][:r], rtol=1e-6)
