"""NumPy emulation of the CUDA pipeline's ALGORITHM -- test infrastructure, CPU only.

It mirrors what the kernels in american_monte_carlo_b200/csrc do (state = cashflow discounted to
time 0, one affine map per column, Hankel moment sums, the lsm_solve.h solver compiled for the host,
Horner evaluation in the internal basis, strict-> exercise test) so that the design can be checked
against the oracle in the build container, where there is no GPU.  The product never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BASIS_ID = {"Power": 0, "Chebyshev": 1, "Legendre": 2, "Laguerre": 3}


def host_solver():
    src = os.path.join(HERE, "native", "solve_host.cpp")
    hdr = os.path.join(ROOT, "american_monte_carlo_b200", "csrc", "lsm_solve.h")
    hdr2 = os.path.join(ROOT, "american_monte_carlo_b200", "csrc", "philox.cuh")
    hdr3 = os.path.join(ROOT, "american_monte_carlo_b200", "csrc", "gbm_quad.cuh")
    out_dir = os.path.join(HERE, "native", "_build")
    so = os.path.join(out_dir, "libamc_solve_host.so")
    os.makedirs(out_dir, exist_ok=True)
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2), os.path.getmtime(hdr3)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.amc_test_lsm_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                       dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       dp, dp, dp, dp, ctypes.POINTER(ctypes.c_int)]
    lib.amc_test_lsm_solve.restype = ctypes.c_int
    return lib


_LIB = None


def solve(degree, basis, scaling, scaling_factor, n_paths, h, g, y_scale, mu_ref, sigma_ref, want_svd=True):
    global _LIB
    if _LIB is None:
        _LIB = host_solver()
    k = degree + 1
    h = np.ascontiguousarray(h, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    gamma = np.zeros(k); beta = np.zeros(k); sv = np.zeros(k); stats = np.zeros(2)
    info = (ctypes.c_int * 3)()
    dp = ctypes.POINTER(ctypes.c_double)
    rc = _LIB.amc_test_lsm_solve(degree, BASIS_ID[basis], int(bool(scaling)), int(bool(want_svd)), float(scaling_factor),
                                 float(n_paths),
                                 h.ctypes.data_as(dp), g.ctypes.data_as(dp), float(y_scale), float(mu_ref),
                                 float(sigma_ref), gamma.ctypes.data_as(dp), beta.ctypes.data_as(dp),
                                 sv.ctypes.data_as(dp), stats.ctypes.data_as(dp), info)
    assert rc == 0
    return dict(gamma=gamma, beta=beta, sv=sv, mean_x=stats[0], std_x=stats[1], rank=info[0], k_internal=info[1],
                sweeps=info[2])


def moments(z, y, degree):
    """h[m] = sum z^m (m <= 2d), g[m] = sum z^m y (m <= d) -- what the streaming kernel reduces."""
    h = np.empty(2 * degree + 1)
    g = np.empty(degree + 1)
    p = np.ones_like(z)
    for m in range(2 * degree + 1):
        h[m] = p.sum()
        if m <= degree:
            g[m] = (p * y).sum()
        p = p * z
    return h, g


def horner(gamma, z):
    acc = np.full_like(z, gamma[-1])
    for c in gamma[-2::-1]:
        acc = acc * z + c
    return acc


def column_maps(paths_tm):
    """Per-column (mu_ref, sigma_ref): shifted one-pass moments, as amc_paths_from_host computes them."""
    n1 = paths_tm.shape[0]
    mu = np.empty(n1); sg = np.empty(n1)
    for t in range(n1):
        col = paths_tm[t]
        c = col[0]
        dlt = col - c
        m1 = dlt.mean(); m2 = (dlt * dlt).mean()
        var = max(m2 - m1 * m1, 0.0)
        mu[t] = c + m1
        s = np.sqrt(var)
        sg[t] = s if s > 1e-14 * max(abs(mu[t]), 1e-300) and s > 0 else 1.0
    return mu, sg


def price(paths, K, r, dt, option_type, barrier_level=None, exercise_type="European", basis_type="Chebyshev",
          degree=4, scaling=False, scaling_factor=2, maps=None, x_dtype=np.float64, want_svd=True):
    """Returns dict(price, tau, U, ranks, gammas).  `paths` is [P, n+1] like the reference's."""
    P, n1 = paths.shape
    n = n1 - 1
    S = np.ascontiguousarray(paths.T.astype(x_dtype)).astype(np.float64)   # time-major, optional f32 rounding
    is_put = option_type == "Put"
    american = exercise_type == "American"
    mu, sg = maps if maps is not None else column_maps(S)
    if barrier_level is not None:
        hit = S <= barrier_level
        first = np.where(hit.any(axis=0), hit.argmax(axis=0), n + 1)          # first knock-in step per path
    else:
        first = np.zeros(P, dtype=np.int64)
    rdt = r * dt

    def intrinsic(x):
        return np.maximum(K - x, 0.0) if is_put else np.maximum(x - K, 0.0)

    U = np.where(first <= n, intrinsic(S[n]) * np.exp(-rdt * n), 0.0)       # discounted to time 0
    tau = np.full(P, n)
    ranks = {}; gammas = {}; sweeps = {}
    for t in range(n - 1, -1, -1):
        z = (S[t] - mu[t]) * (1.0 / sg[t])
        h, g = moments(z, U, degree)
        res = solve(degree, basis_type, scaling, scaling_factor, P, h, g, np.exp(rdt * t), mu[t], sg[t], want_svd)
        ranks[t] = res["rank"]; gammas[t] = res["gamma"]; sweeps[t] = res["sweeps"]
        if american:
            fit = horner(res["gamma"], z)
            iv = intrinsic(S[t])
            take = (first <= t) & (iv > 0) & (iv > fit)
            U = np.where(take, iv * np.exp(-rdt * t), U)
            tau = np.where(take, t, tau)
    return dict(price=U.sum() / P, tau=tau, U=U, ranks=ranks, gammas=gammas, sweeps=sweeps)
