"""CPU oracle for the Longstaff-Schwartz hot path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference algorithm
(`/root/reference/american_monte_carlo.py`, lines 72-197; "amc.py" below).  It exists so
that the CUDA path can be checked against the reference's arithmetic on a GPU box where
`/root/reference` does not exist.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
(`american_monte_carlo_b200`) never does and has no CPU path of its own.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference in
the build container (matplotlib/QuantLib stubbed, they are not on the hot path), checks that
every function below returns bit-identical results to it, and writes
`tests/golden/golden.json` (notebook prices `AmericanMonteCarlo.ipynb:202-203,248-249,
312-313,377-378,442-443`, the 12 `unit_test.py:30-41` cases, config 1/2 prices, per-step
lstsq rank / singular values / coefficients).  `tests/test_oracle_golden.py` re-checks the
oracle against that file on every run.

The arithmetic that lives in a third-party dependency is NumPy's: the legacy global
MT19937 + polar-Gauss stream (`np.random.normal`) and `np.linalg.lstsq(rcond=None)`
(LAPACK gelsd: truncated SVD, singular values <= eps*max(M,N)*s_max treated as zero).
Both are called here exactly as the reference calls them, so the oracle inherits their
behaviour instead of restating it.  The reference pins no NumPy version; the goldens were
produced with numpy 2.3.5 / OpenBLAS 0.3.30.

Expression order is kept identical to the reference wherever floating-point rounding could
differ (e.g. `-r * dt * (tau - t)`), because parity is asserted bitwise against it.
"""
from __future__ import annotations

import numpy as np

BASES = ("Power", "Chebyshev", "Legendre")
# "Laguerre" is NOT in the reference (amc.py:99-101 has three bases).  BASELINE.json's
# config 5 names it, so the oracle carries it as a clearly-marked extension evaluated the
# same way the reference evaluates the other orthogonal families (numpy.polynomial, one-hot
# coefficient vector, no domain mapping).
EXTENSION_BASES = ("Laguerre",)


# --------------------------------------------------------------------------- path simulation
def draw_normals(n_paths: int, n_time_steps: int) -> np.ndarray:
    """The reference's draw: amc.py:74 (legacy global RNG, row-major fill [P, n])."""
    return np.random.normal(size=(n_paths, n_time_steps))


def paths_from_normals(Z: np.ndarray, S0, r, sigma, T) -> np.ndarray:
    """GBM exact discretisation from given normals: amc.py:73,75-81."""
    n_paths, n_time_steps = Z.shape
    dt = T / n_time_steps
    log_step = (r - 0.5 * sigma ** 2) * dt + sigma * np.sqrt(dt) * Z          # amc.py:75
    step_factor = np.exp(log_step)                                            # amc.py:76
    out = np.zeros((n_paths, n_time_steps + 1))                               # amc.py:78
    out[:, 0] = S0                                                            # amc.py:79
    out[:, 1:] = S0 * np.cumprod(step_factor, axis=1)                         # amc.py:80
    return out


def generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths) -> np.ndarray:
    """amc.py:72-81."""
    return paths_from_normals(draw_normals(n_paths, n_time_steps), S0, r, sigma, T)


# --------------------------------------------------------------------------- payoff / barrier
def payoff(S, K, option_type="Call"):
    """amc.py:85-86 -- anything that is not the string "Put" is a call."""
    if option_type == "Put":
        return np.maximum(K - S, 0)
    return np.maximum(S - K, 0)


def knock_in_flags(paths: np.ndarray, barrier_level) -> np.ndarray:
    """Down-and-in flags, running OR over time: amc.py:171-176 (`is not None` test)."""
    if barrier_level is None:
        return np.ones_like(paths, dtype=bool)                                # amc.py:175
    return np.maximum.accumulate(paths <= barrier_level, axis=1)             # amc.py:173


# --------------------------------------------------------------------------- regression
def basis_matrix(X: np.ndarray, basis_type: str, degree: int) -> np.ndarray:
    """Design matrix [len(X), degree+1], no domain mapping: amc.py:98-106."""
    if basis_type == "Power":
        cols = [X ** i for i in range(degree + 1)]                            # amc.py:99
    elif basis_type == "Chebyshev":
        cols = [np.polynomial.chebyshev.chebval(X, [0] * i + [1]) for i in range(degree + 1)]
    elif basis_type == "Legendre":
        cols = [np.polynomial.legendre.legval(X, [0] * i + [1]) for i in range(degree + 1)]
    elif basis_type == "Laguerre":                                            # extension, see top
        cols = [np.polynomial.laguerre.lagval(X, [0] * i + [1]) for i in range(degree + 1)]
    else:                                                                     # amc.py:103-104
        raise ValueError(f"Unknown basis type '{basis_type}'. Use 'Power', 'Chebyshev', or 'Legendre'.")
    return np.column_stack(cols)                                              # amc.py:106


def regression_fit(X, Y, basis_type="Power", degree=3, scaling=False, scaling_factor=2, diag=None):
    """Fitted values of the least-squares regression: amc.py:110-122.

    `diag`, when a dict, receives numpy's own rank / singular values / coefficients and the
    standardisation constants, for the golden file and for diagnosing truncation decisions.
    """
    if scaling:
        centre = np.mean(X)                                                   # amc.py:112
        spread = max(np.std(X), 1e-6)                                         # amc.py:113
        U = (X - centre) / (scaling_factor * spread)                          # amc.py:114
    else:
        centre, spread, U = 0.0, 1.0, X
    A = basis_matrix(U, basis_type, degree)                                   # amc.py:116/120
    coeffs, _, rank, sv = np.linalg.lstsq(A, Y, rcond=None)                   # amc.py:117/121
    if diag is not None:
        diag.update(rank=int(rank), sv=np.array(sv), coeffs=np.array(coeffs),
                    centre=float(centre), spread=float(spread))
    return A @ coeffs                                                         # amc.py:118/122


def continuation_estimate(paths, t, r, dt, cashflows, exercise_times, basis_type, degree,
                          diag=None, **kwargs):
    """amc.py:126-135: regress discounted future cashflows of ALL paths on S_t, clamp at 0."""
    X = paths[:, t]                                                           # amc.py:127
    Y = cashflows * np.exp(-r * dt * (exercise_times - t))                    # amc.py:128
    if len(X) > 0:                                                            # amc.py:130
        return np.maximum(regression_fit(X, Y, basis_type, degree, diag=diag, **kwargs), 0)
    return np.zeros(paths.shape[0])                                           # amc.py:134


# --------------------------------------------------------------------------- LSM driver
class LsmResult:
    """Everything a parity test wants to look at (the reference returns only the first two)."""
    __slots__ = ("price", "continuation_values", "cashflows", "exercise_times", "steps")

    def __init__(self, price, continuation_values, cashflows, exercise_times, steps):
        self.price = price
        self.continuation_values = continuation_values
        self.cashflows = cashflows
        self.exercise_times = exercise_times
        self.steps = steps          # {t: diag dict} for t = 0..n-1


def lsm_backward(paths, K, r, dt, option_type, barrier_level=None, exercise_type="European",
                 basis_type="Chebyshev", degree=4, keep_continuation=True, keep_diag=False,
                 **kwargs) -> LsmResult:
    """amc.py:139-167 (backward sweep) + amc.py:180-197 (state set-up and final mean).

    `keep_continuation=False` skips the two per-step [P] copies (amc.py:164), which the
    reference always makes; it does not change any number.
    """
    n_paths, n_cols = paths.shape                                             # amc.py:184
    n = n_cols - 1
    cashflows = np.zeros(n_paths)                                             # amc.py:186
    exercise_times = np.full(n_paths, n)                                      # amc.py:187
    hit = knock_in_flags(paths, barrier_level)                                # amc.py:189
    cont_list = []
    steps = {}

    for t in range(n, -1, -1):                                                # amc.py:141
        hit_t = hit[:, t]
        cont = np.zeros(n_paths)                                              # amc.py:145
        if t == n:                                                            # amc.py:147-149
            cashflows[hit_t] = payoff(paths[hit_t, t], K, option_type)
            exercise_times[hit_t] = t
        else:
            diag = {} if keep_diag else None
            cont = continuation_estimate(paths, t, r, dt, cashflows, exercise_times,
                                         basis_type, degree, diag=diag, **kwargs)   # amc.py:151
            if keep_diag:
                steps[t] = diag
            if exercise_type == 'American':                                   # amc.py:154
                now = payoff(paths[:, t], K, option_type)
                # amc.py:155-162 and 90-94: candidates are knocked-in AND in the money;
                # they exercise when the immediate payoff is STRICTLY above the estimate.
                take = hit_t & (now > 0) & (now > cont)
                cashflows[take] = now[take]
                exercise_times[take] = t
        if keep_continuation:
            cont_list.append((t, paths[:, t].copy(), cont.copy()))            # amc.py:164
    cont_list.reverse()                                                       # amc.py:167

    price = np.mean(cashflows * np.exp(-r * dt * exercise_times))             # amc.py:196
    return LsmResult(price, cont_list, cashflows, exercise_times, steps)


def lsmc_option_pricing(paths, K, r, dt, option_type, barrier_level=None,
                        exercise_type="European", basis_type="Chebyshev", degree=4, **kwargs):
    """Same signature and return value as amc.py:180-197."""
    res = lsm_backward(paths, K, r, dt, option_type, barrier_level, exercise_type,
                       basis_type, degree, **kwargs)
    return res.price, res.continuation_values


# --------------------------------------------------------------------------- CCR exposures
def ccr_exposures(continuation_values):
    """amc.py:400-414: per step (t, 5th pct, 95th pct, mean) of the finite continuation values."""
    out = []
    for t, _, cont in continuation_values:
        ok = cont[np.isfinite(cont)]
        if len(ok) == 0:
            out.append((t, np.nan, np.nan, np.nan))
        else:
            out.append((t, np.percentile(ok, 5), np.percentile(ok, 95), np.mean(ok)))
    return out
