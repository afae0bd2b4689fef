"""CPU tier: american_monte_carlo_b200/csrc/lsm_solve.h compiled for the host, against numpy.linalg.lstsq.

The solver is the numerically delicate part of the CUDA path (it must reproduce numpy's rank-truncated SVD fit
from Gram sums); this runs the device source on the host so it can be checked without a GPU.
"""
import numpy as np
import pytest

import pipeline_emulator as emu
from conftest import oracle_case
from oracle import lsm_oracle as orc


def _fit_via_solver(X, Y, basis, degree, scaling=False, scaling_factor=2):
    mu, sg = X.mean(), X.std()
    sg = sg if sg > 0 else 1.0
    z = (X - mu) * (1.0 / sg)
    h, g = emu.moments(z, Y, degree)
    res = emu.solve(degree, basis, scaling, scaling_factor, len(X), h, g, 1.0, mu, 1.0 / (1.0 / sg))
    return emu.horner(res["gamma"], z), res


@pytest.mark.parametrize("basis", ["Power", "Chebyshev", "Legendre", "Laguerre"])
@pytest.mark.parametrize("degree", [0, 1, 2, 3, 5, 8, 10])
@pytest.mark.parametrize("scaling", [False, True])
def test_fitted_values_match_numpy_lstsq(basis, degree, scaling):
    rng = np.random.default_rng(degree * 7 + len(basis))
    X = 36.0 * np.exp(0.2 * rng.standard_normal(20000) - 0.02)
    Y = np.maximum(40.0 - X, 0) * np.exp(0.1 * rng.standard_normal(X.size))
    diag = {}
    want = orc.regression_fit(X, Y, basis, degree, scaling=scaling, scaling_factor=2, diag=diag)
    got, res = _fit_via_solver(X, Y, basis, degree, scaling, 2)
    assert res["rank"] == diag["rank"]
    scale = np.abs(want).max()
    # (a) against the exact truncated projection U_r U_r^T Y of the same design matrix (backward-stable SVD)
    U_ = (X - diag["centre"]) / (2 * diag["spread"]) if scaling else X
    A = orc.basis_matrix(U_, basis, degree)
    Us, ss, _ = np.linalg.svd(A, full_matrices=False)
    r = diag["rank"]
    cond_kept = ss[0] / ss[r - 1]
    exact = Us[:, :r] @ (Us[:, :r].T @ Y)
    assert np.abs(got - exact).max() <= max(1e-9, 4 * cond_kept * 2.2e-16) * scale
    # (b) against numpy's lstsq itself, whose own fitted values carry an error ~ cond(A) * eps (gelsd): at
    #     cond 7e10 (unscaled Power-5) lstsq is 4e-8 away from the exact projection while this solver is 3e-11 away
    assert np.abs(got - want).max() <= max(2e-9, 50 * cond_kept * 2.2e-16) * scale
    k = degree + 1
    np.testing.assert_allclose(res["sv"][:res["rank"]], diag["sv"][:res["rank"]], rtol=1e-6)
    if diag["rank"] == k and np.linalg.cond(orc.basis_matrix((X - diag["centre"]) / (2 * diag["spread"]) if scaling else X,
                                                             basis, degree)) < 1e6:
        np.testing.assert_allclose(res["beta"], diag["coeffs"], rtol=1e-5, atol=1e-9 * np.abs(diag["coeffs"]).max())


def test_constant_column_gives_mean_of_y():
    # t = 0 of every sweep: all paths at S0, numpy's min-norm solution fits mean(Y)
    X = np.full(1000, 36.0)
    Y = np.random.default_rng(0).random(1000)
    want = orc.regression_fit(X, Y, "Power", 3)
    z = (X - 36.5) / 2.0
    h, g = emu.moments(z, Y, 3)
    res = emu.solve(3, "Power", False, 2, 1000, h, g, 1.0, 36.5, 2.0)
    assert res["rank"] == 1 and res["k_internal"] == 1
    np.testing.assert_allclose(emu.horner(res["gamma"], z), want, rtol=1e-13)
    np.testing.assert_allclose(res["gamma"][0], Y.mean(), rtol=1e-14)


def test_y_scale_multiplies_the_fit():
    rng = np.random.default_rng(3)
    X = rng.normal(100, 15, 5000)
    Y = rng.random(5000)
    z = (X - 100) / 15
    h, g = emu.moments(z, Y, 4)
    a = emu.solve(4, "Chebyshev", True, 2, 5000, h, g, 1.0, 100.0, 15.0)
    b = emu.solve(4, "Chebyshev", True, 2, 5000, h, g, 1.25, 100.0, 15.0)
    np.testing.assert_allclose(b["gamma"], 1.25 * a["gamma"], rtol=1e-14)
    np.testing.assert_allclose(a["mean_x"], X.mean(), rtol=1e-13)
    np.testing.assert_allclose(a["std_x"], X.std(), rtol=1e-12)


def test_rank_rule_uses_global_path_count():
    # the same sums with a larger GLOBAL path count must truncate more (numpy: rcond = eps * max(P, k))
    rng = np.random.default_rng(5)
    X = 36.0 * np.exp(0.028 * rng.standard_normal(50000))       # a t=1-like narrow column
    Y = np.maximum(40 - X, 0)
    mu, sg = X.mean(), X.std()
    z = (X - mu) / sg
    h, g = emu.moments(z, Y, 3)
    r_small = emu.solve(3, "Power", False, 2, 50000, h, g, 1.0, mu, sg)["rank"]
    r_big = emu.solve(3, "Power", False, 2, 5e9, h * 1e5, g * 1e5, 1.0, mu, sg)["rank"]
    assert r_small == np.linalg.lstsq(orc.basis_matrix(X, "Power", 3), Y, rcond=None)[2]
    assert r_big < r_small


@pytest.mark.parametrize("name", ["nb_american_put", "ut_Put_American_None", "ut_Call_American_80", "c1_power3",
                                  "c1_di30_power3", "small_power8_unscaled", "small_legendre8_scaled",
                                  "small_chebyshev10_unscaled", "small_degree0", "deep_itm_exercise_at_0",
                                  "tiny_paths_lt_k"])
@pytest.mark.parametrize("want_svd", [True, False])
def test_emulated_pipeline_matches_oracle_with_zero_flips(golden, name, want_svd):
    """The whole algorithm the kernels implement (time-0 discounted state, per-column maps, Hankel sums, this solver,
    Horner decisions) against the oracle: same price to 1e-12 and not one path exercising at a different step."""
    c = golden[name]
    _, paths, o = oracle_case(c)
    dt = c["T"] / c["n_time_steps"]
    e = emu.price(paths, c["K"], c["r"], dt, c["option_type"], c["barrier_level"], c["exercise_type"], c["basis_type"],
                  c["degree"], want_svd=want_svd, **c["kwargs"])
    if not want_svd and name in ("c1_power3", "small_legendre8_scaled", "nb_american_put"):
        # the full-rank certificate must actually skip the SVD on most steps of well-conditioned sweeps
        skipped = sum(1 for v in e["sweeps"].values() if v == -1)
        assert skipped >= 0.8 * c["n_time_steps"], skipped
    assert int((e["tau"] != o.exercise_times).sum()) == 0
    assert [e["ranks"][t] for t in range(c["n_time_steps"])] == c["ranks"]
    assert abs(e["price"] - c["price"]) <= 1e-12 * max(abs(c["price"]), 1.0)


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10, through the host build of csrc/philox.cuh."""
    lib = emu.host_solver()
    import ctypes
    out = (ctypes.c_uint32 * 4)()
    lib.amc_test_philox.argtypes = [ctypes.c_uint32] * 6 + [ctypes.POINTER(ctypes.c_uint32)]
    lib.amc_test_philox(0, 0, 0, 0, 0, 0, out)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = 0xFFFFFFFF
    lib.amc_test_philox(f, f, f, f, f, f, out)
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    lib.amc_test_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert [hex(v) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_fixed_point_scale_of_the_float_generator():
    """gbm_quad.cuh::fixed_point_bits (host build): the log2-price of the float generator is an int32 sum of increments in
    units of 2^-k.  k must keep the largest possible increment (Box-Muller radius <= 6.8) inside the exact range of the
    magic-number float -> int conversion (2^22) and an 8-sigma excursion of the whole path inside int32, over the whole
    range of markets and step counts -- including sigma = 0, where only the drift sets the scale."""
    import ctypes
    import math
    lib = emu.host_solver()
    lib.amc_test_fixed_point_bits.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.amc_test_fixed_point_bits.restype = ctypes.c_int
    log2e = 1.4426950408889634
    seen = set()
    for sigma in (0.0, 0.01, 0.2, 0.6, 1.5):
        for T in (0.01, 1.0, 30.0):
            for n in (1, 4, 50, 252, 5000):
                for r in (0.0, 0.06, -0.02):
                    dt = T / n
                    d2 = (r - 0.5 * sigma * sigma) * dt * log2e
                    v2 = sigma * math.sqrt(dt) * log2e
                    k = lib.amc_test_fixed_point_bits(d2, v2, n)
                    seen.add(k)
                    assert 4 <= k <= 30
                    gmax = abs(d2) + v2 * 6.8
                    span = n * abs(d2) + v2 * (8.0 * math.sqrt(n) + 7.0) + 1.0
                    if k > 4:                                   # k = 4 is the floor for absurd markets
                        assert gmax * 2.0 ** k <= 2.0 ** 22 * (1 + 1e-12), (sigma, T, n, r, k)
                        assert span * 2.0 ** k <= 2.0 ** 30 * (1 + 1e-12), (sigma, T, n, r, k)
                    # not wastefully small: one more bit would break one of the two bounds (or hit the cap)
                    assert k == 30 or gmax * 2.0 ** (k + 1) > 2.0 ** 22 or span * 2.0 ** (k + 1) > 2.0 ** 30
    assert lib.amc_test_fixed_point_bits(0.06 / 252 * log2e - 0.02 / 252 * log2e, 0.2 * math.sqrt(1 / 252) * log2e, 252) == 25
    assert len(seen) > 5
