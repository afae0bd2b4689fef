"""GPU tier: the device generators themselves.

* the integer Philox stage as compiled for the DEVICE (philox.cuh takes a different mulhilo branch under
  __CUDA_ARCH__) against the Random123 known-answer vectors and the host build of the same header -- bit-exact;
* the float Box-Muller normals exactly as the path kernel forms them (MUFU lg2 / sqrt / sin / cos approximations):
  1e8 samples, Kolmogorov-Smirnov distance on a 4096-bin CDF, tail mass beyond 4 sigma within 3 standard errors,
  |z| beyond 5.5 reached, first four moments.
"""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KAT = [  # Random123 kat_vectors, philox4x32-10: counter, key, expected
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _device_philox(amc, rounds, counters, key):
    from american_monte_carlo_b200 import _native as N
    ctr = np.ascontiguousarray(counters, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.empty_like(ctr)
    N.check(N.lib().amc_selftest_philox(amc.default_context().handle, rounds, ctr.ctypes.data, k.ctypes.data, len(ctr),
                                        out.ctypes.data))
    return out


def test_device_philox_known_answer_vectors(amc):
    for ctr, key, want in KAT:
        got = _device_philox(amc, 10, [ctr], key)[0]
        assert [hex(v) for v in got] == [hex(v) for v in want]


def test_device_philox_equals_host_build_on_random_counters(amc):
    import pipeline_emulator as emu
    host = emu.host_solver()
    host.amc_test_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    rng = np.random.default_rng(3)
    ctr = rng.integers(0, 2 ** 32, size=(2000, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2 ** 32, size=2, dtype=np.uint64).astype(np.uint32)
    got = _device_philox(amc, 10, ctr, key)
    out = (C.c_uint32 * 4)()
    for i in range(len(ctr)):
        host.amc_test_philox(*[int(v) for v in ctr[i]], int(key[0]), int(key[1]), out)
        assert list(out) == got[i].tolist(), i
    # the 7-round option is a different (documented) generator, not a truncation bug: it must differ from 10 rounds
    assert not np.array_equal(_device_philox(amc, 7, ctr[:8], key), got[:8])


@pytest.mark.parametrize("rounds", [10, 7])
def test_float_box_muller_distribution_1e8(amc, rounds):
    from american_monte_carlo_b200 import _native as N
    n_quads, n_steps, bins, lo, hi = 1_000_000, 25, 4096, -8.0, 8.0
    hist = np.zeros(bins + 2, dtype=np.uint64)
    stats = np.zeros(6)
    N.check(N.lib().amc_selftest_normals(amc.default_context().handle, rounds, C.c_uint64(20261018), n_quads, n_steps, bins,
                                         lo, hi, hist.ctypes.data, stats.ctypes.data))
    n = 4 * n_quads * n_steps
    assert int(hist.sum()) == n == int(stats[0]) == 100_000_000
    assert hist[0] == 0 and hist[-1] == 0                       # |z| <= 6.76 by construction (u1 >= 2^-33)
    edges = lo + (hi - lo) * np.arange(1, bins + 1) / bins      # right edges of the regular bins
    cdf_emp = np.cumsum(hist[1:-1].astype(np.float64)) / n
    Phi = np.array([0.5 * math.erfc(-e / math.sqrt(2.0)) for e in edges])
    ks = float(np.max(np.abs(cdf_emp - Phi)))
    assert ks < 1.63 / math.sqrt(n), ks                          # 1 % critical value of the KS statistic
    # tails: mass beyond 4 sigma (each side) within 3 standard errors of 3.167e-5
    p4 = 0.5 * math.erfc(4.0 / math.sqrt(2.0))
    k4 = int(round((4.0 - lo) / (hi - lo) * bins))
    upper = float(hist[1 + k4:-1].sum())
    lower = float(hist[1:1 + bins - k4].sum())
    se = math.sqrt(n * p4)
    assert abs(upper - n * p4) < 3 * se and abs(lower - n * p4) < 3 * se, (upper, lower, n * p4, se)
    assert stats[5] >= 5.5                                       # the far tail is reached (expected max ~5.7 at 1e8)
    m1, m2, m3, m4 = (stats[i] / n for i in (1, 2, 3, 4))
    assert abs(m1) < 5 / math.sqrt(n) and abs(m2 - 1) < 5 * math.sqrt(2 / n)
    assert abs(m3) < 5 * math.sqrt(15 / n) and abs(m4 - 3) < 5 * math.sqrt(96 / n)
