"""QuantLib-free benchmark prices: a stand-in for `get_quantlib_option` (amc.py:57-68).

NOT part of the hot path and not accelerated: the reference uses QuantLib only as an independent yardstick for the
Monte Carlo price (unit_test.py:15-21, tolerance 0.2).  QuantLib is not installed in this image, so the same four
engines the reference wires up (amc.py:26-46) are restated from their textbook definitions:
  European vanilla        AnalyticEuropeanEngine        -> Black-Scholes-Merton closed form
  American vanilla        BinomialVanillaEngine("crr")  -> Cox-Ross-Rubinstein tree, n_steps
  European down-and-in    AnalyticBarrierEngine         -> Reiner-Rubinstein (1991) formulas, rebate 0 (amc.py:63-64)
  American down-and-in    BinomialBarrierEngine("crr")  -> CRR tree; a node at or below the barrier is worth the American
                                                           vanilla option of the same tree
They reproduce the three QuantLib values stored in the reference notebook to the printed 4 d.p. (9.8928, 10.0198,
4.0316; tests/test_benchmarks.py).  Maturity follows amc.py:27: `today + int(T*365)` days under Actual/365.
"""
from __future__ import annotations

import math


def _ncdf(x):
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def _maturity(T):
    return int(T * 365) / 365.0                                       # amc.py:27,32,39,42


def black_scholes(S, K, r, T, sigma, is_put, q=0.0):
    if T <= 0:
        return max(K - S, 0.0) if is_put else max(S - K, 0.0)
    sd = sigma * math.sqrt(T)
    d1 = (math.log(S / K) + (r - q + 0.5 * sigma * sigma) * T) / sd
    d2 = d1 - sd
    if is_put:
        return K * math.exp(-r * T) * _ncdf(-d2) - S * math.exp(-q * T) * _ncdf(-d1)
    return S * math.exp(-q * T) * _ncdf(d1) - K * math.exp(-r * T) * _ncdf(d2)


def _crr_params(r, q, sigma, T, n):
    dt = T / n
    u = math.exp(sigma * math.sqrt(dt))
    d = 1.0 / u
    p = (math.exp((r - q) * dt) - d) / (u - d)
    return dt, u, d, p, math.exp(-r * dt)


def crr_american(S, K, r, T, sigma, n, is_put, q=0.0, barrier=None):
    """CRR tree.  With `barrier` (down-and-in): no value and no exercise until a node at or below the barrier is
    reached, where the contract becomes the American vanilla option."""
    if T <= 0:
        return max(K - S, 0.0) if is_put else max(S - K, 0.0)
    dt, u, d, p, disc = _crr_params(r, q, sigma, T, n)

    def payoff(x):
        return max(K - x, 0.0) if is_put else max(x - K, 0.0)

    spot = [S * u ** j * d ** (n - j) for j in range(n + 1)]
    van = [payoff(x) for x in spot]
    bar = [v if (barrier is not None and x <= barrier) else 0.0 for v, x in zip(van, spot)] if barrier is not None else None
    for i in range(n - 1, -1, -1):
        spot = [S * u ** j * d ** (i - j) for j in range(i + 1)]
        van = [max(disc * (p * van[j + 1] + (1 - p) * van[j]), payoff(spot[j])) for j in range(i + 1)]
        if bar is not None:
            bar = [van[j] if spot[j] <= barrier else disc * (p * bar[j + 1] + (1 - p) * bar[j]) for j in range(i + 1)]
    return bar[0] if bar is not None else van[0]


def down_and_in_european(S, K, r, T, sigma, H, is_put, q=0.0):
    """Reiner-Rubinstein, rebate 0.  Raises RuntimeError when the barrier is already touched, like QuantLib's
    AnalyticBarrierEngine (the reference relies on that at amc.py:219)."""
    if S <= H:
        raise RuntimeError("barrier touched")
    b = r - q
    sd = sigma * math.sqrt(T)
    mu = (b - 0.5 * sigma * sigma) / (sigma * sigma)
    x1 = math.log(S / K) / sd + (1 + mu) * sd
    x2 = math.log(S / H) / sd + (1 + mu) * sd
    y1 = math.log(H * H / (S * K)) / sd + (1 + mu) * sd
    y2 = math.log(H / S) / sd + (1 + mu) * sd
    phi = -1.0 if is_put else 1.0
    eta = 1.0
    cS, cK = S * math.exp((b - r) * T), K * math.exp(-r * T)
    A = phi * cS * _ncdf(phi * x1) - phi * cK * _ncdf(phi * x1 - phi * sd)
    B = phi * cS * _ncdf(phi * x2) - phi * cK * _ncdf(phi * x2 - phi * sd)
    Cc = phi * cS * (H / S) ** (2 * (mu + 1)) * _ncdf(eta * y1) - phi * cK * (H / S) ** (2 * mu) * _ncdf(eta * y1 - eta * sd)
    Dd = phi * cS * (H / S) ** (2 * (mu + 1)) * _ncdf(eta * y2) - phi * cK * (H / S) ** (2 * mu) * _ncdf(eta * y2 - eta * sd)
    if is_put:
        return B - Cc + Dd if K > H else A
    return Cc if K > H else A - B + Dd


class BenchmarkOption:
    """Object with the one method the reference's callers use: `.NPV()` (unit_test.py:16, amc.py:217,501)."""

    def __init__(self, fn):
        self._fn = fn

    def NPV(self):
        return self._fn()


def get_quantlib_option(S0, K, r, T, sigma, n_steps=100, option_type="Call", exercise_type="European",
                        barrier_level=None, dividend_yield=0.0):
    """Same signature as amc.py:57-58; returns an object with `.NPV()`."""
    is_put = option_type == "Put"                                     # amc.py:60
    Tm = _maturity(T)
    if barrier_level is not None:                                     # amc.py:37-46
        if exercise_type == "European":
            return BenchmarkOption(lambda: down_and_in_european(S0, K, r, Tm, sigma, barrier_level, is_put, dividend_yield))
        if exercise_type == "American":
            return BenchmarkOption(lambda: crr_american(S0, K, r, Tm, sigma, n_steps, is_put, dividend_yield,
                                                        barrier=barrier_level))
        raise NotImplementedError("Barrier options with this exercise type are not implemented.")   # amc.py:45
    engines = {                                                       # amc.py:48-53 (KeyError on anything else)
        "European": lambda: black_scholes(S0, K, r, Tm, sigma, is_put, dividend_yield),
        "American": lambda: crr_american(S0, K, r, Tm, sigma, n_steps, is_put, dividend_yield),
    }
    return BenchmarkOption(engines[exercise_type])
