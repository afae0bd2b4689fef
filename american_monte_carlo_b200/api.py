"""Host side of the drop-in: the reference's Python entry points on top of libamc (CUDA, sm_100a).

Mirrors the function names, argument order/meaning and error behaviour of
/root/reference/american_monte_carlo.py:72-197 ("amc.py").  Every function computes on the GPU through the
C ABI of include/amc.h; nothing here has a NumPy/CPU implementation and a missing library or device raises.

What differs from the reference, by design:
  * `generate_asset_paths` returns a `DevicePaths` (device-resident, timestep-major) instead of an ndarray.  It has
    `.shape == (n_paths, n_time_steps + 1)`, `__array__`, row indexing and `len()` -- everything the reference's
    consumers touch (amc.py:184,206; plots.py:12) -- and is accepted back by `lsmc_option_pricing`.
  * `lsmc_option_pricing` returns `(price, ContinuationValues)`: the second element is a lazy sequence of the same
    `(t, S_t, continuation_t)` tuples (amc.py:164) that only materialises a step when it is indexed.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections.abc import Sequence

import numpy as np

from . import _native as N

_BASES = ("Power", "Chebyshev", "Legendre")          # amc.py:99-101
# conditioning report of the solve (include/amc.h: amc_lsm_steps.pivot_loss): parity with lstsq is tested up to here
PIVOT_LOSS_WARN = 1e12
_EXTRA_BASES = ("Laguerre",)                         # addition, BASELINE.json config 5


# --------------------------------------------------------------------------------------------- context
class Context:
    """One per process and GPU.  Owns the libamc context (stream, scratch memory, optional NCCL communicator)."""

    def __init__(self, device=None, stream=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)
        h = C.c_void_p()
        N.check(N.lib().amc_ctx_create(self.device, C.c_void_p(stream or 0), C.byref(h)))
        self.handle = h
        self.world_size, self.rank = 1, 0

    def close(self):
        if self.handle:
            N.lib().amc_ctx_destroy(self.handle)
            self.handle = None

    def sync(self):
        N.check(N.lib().amc_ctx_sync(self.handle))

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        N.check(N.lib().amc_ctx_device_info(self.handle, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), total_mem=mem.value)

    # multi-GPU: the 128-byte NCCL id travels over whatever the caller uses for rendezvous (torch.distributed)
    @staticmethod
    def new_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        N.check(N.lib().amc_comm_unique_id(buf))
        return buf.raw

    def init_comm(self, world_size, rank, unique_id: bytes):
        N.check(N.lib().amc_comm_init(self.handle, int(world_size), int(rank), C.c_char_p(unique_id)))
        self.world_size, self.rank = int(world_size), int(rank)

    @property
    def transport(self) -> str:
        """How the per-step all-reduce travels: 'none' (one GPU), 'nccl', or 'p2p' (fused, NVLink peer memory)."""
        t = C.c_int()
        N.check(N.lib().amc_comm_transport(self.handle, C.byref(t)))
        return {0: "none", 1: "nccl", 2: "p2p"}[t.value]

    def allreduce_host(self, arr: np.ndarray) -> np.ndarray:
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        N.check(N.lib().amc_comm_allreduce_host(self.handle, arr.ctypes.data_as(N.c_double_p), arr.size))
        return arr


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def set_default_context(ctx: Context | None):
    global _default_ctx
    _default_ctx = ctx


# --------------------------------------------------------------------------------------------- paths
_DTYPES = {"float64": N.F64, "f64": N.F64, np.float64: N.F64, "float32": N.F32, "f32": N.F32, np.float32: N.F32}


def _dtype_id(dtype):
    try:
        return _DTYPES[dtype]
    except (KeyError, TypeError):
        d = np.dtype(dtype)
        if d == np.float64:
            return N.F64
        if d == np.float32:
            return N.F32
        raise ValueError(f"unsupported path dtype {dtype!r}: use float64 or float32")


class DevicePaths:
    """A path set resident in HBM, timestep-major.  Stands in for the [n_paths, n_time_steps+1] ndarray of amc.py:78-81."""

    def __init__(self, ctx: Context, handle, n_paths_local, n_paths_global, n_time_steps, dtype_id, path_offset=0):
        self.ctx, self.handle = ctx, handle
        self.n_paths_local, self.n_paths_global = int(n_paths_local), int(n_paths_global)
        self.n_time_steps, self.dtype_id, self.path_offset = int(n_time_steps), dtype_id, int(path_offset)

    # --- what the reference's consumers use
    @property
    def shape(self):
        return (self.n_paths_local, self.n_time_steps + 1)

    @property
    def dtype(self):
        return np.dtype(np.float64 if self.dtype_id == N.F64 else np.float32)

    @property
    def ndim(self):
        return 2

    def __len__(self):
        return self.n_paths_local

    def rows(self, p0, p1) -> np.ndarray:
        out = np.empty((p1 - p0, self.n_time_steps + 1), dtype=np.float64)
        N.check(N.lib().amc_paths_rows(self.handle, p0, p1, out.ctypes.data))
        return out

    def column(self, t) -> np.ndarray:
        if t < 0:
            t += self.n_time_steps + 1
        out = np.empty(self.n_paths_local, dtype=np.float64)
        N.check(N.lib().amc_paths_column(self.handle, int(t), out.ctypes.data))
        return out

    def __array__(self, dtype=None, copy=None):
        out = self.rows(0, self.n_paths_local)
        return out if dtype is None else out.astype(dtype, copy=False)

    def __getitem__(self, key):
        # rows for plotting (plots.py:12 `paths[i]`, amc.py:206 `paths[:n]`); [:, t] columns (amc.py:127)
        if isinstance(key, tuple) and len(key) == 2 and isinstance(key[0], slice) and key[0] == slice(None) \
                and isinstance(key[1], (int, np.integer)):
            return self.column(int(key[1]))
        if isinstance(key, (int, np.integer)):
            k = int(key) + (self.n_paths_local if key < 0 else 0)
            if not 0 <= k < self.n_paths_local:
                raise IndexError(key)
            return self.rows(k, k + 1)[0]
        if isinstance(key, slice):
            start, stop, step = key.indices(self.n_paths_local)
            if step == 1:
                return self.rows(start, max(stop, start))
        return np.asarray(self)[key]

    def column_maps(self):
        mu = np.empty(self.n_time_steps + 1)
        sg = np.empty(self.n_time_steps + 1)
        N.check(N.lib().amc_paths_column_maps(self.handle, mu.ctypes.data, sg.ctypes.data))
        return mu, sg

    @property
    def nbytes_device(self):
        b = C.c_int64()
        N.check(N.lib().amc_paths_info(self.handle, None, None, None, None, C.byref(b)))
        return b.value

    def free(self):
        if self.handle:
            N.lib().amc_paths_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def shard_range(n_paths_global: int, world_size: int, rank: int, unit: int = 1):
    """Contiguous block partition of the path axis: rank g owns [lo, hi).  Sizes differ by at most `unit` paths; with
    unit=4 every shard starts on a quad boundary (one Philox call of the float generator serves four adjacent paths:
    aligned shards keep 128-bit stores and are what the path-free mode requires)."""
    n, unit = int(n_paths_global), int(unit)
    blocks = -(-n // unit)
    base, rem = divmod(blocks, int(world_size))
    lo_b = rank * base + min(rank, rem)
    hi_b = lo_b + base + (1 if rank < rem else 0)
    return min(lo_b * unit, n), min(hi_b * unit, n)


def generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths, *, rng="numpy", dtype="float64", seed=None,
                         store_paths=True, ctx: Context | None = None):
    """amc.py:72-81 on the GPU.  `n_paths` is the GLOBAL path count; under a multi-rank context each rank
    simulates its own contiguous shard.

    rng="numpy"  (default, reference-compatible): the standard normals are drawn on the host from NumPy's legacy
                 global stream exactly as amc.py:74 does -- np.random.seed(...) by the caller gives the reference's
                 paths -- and the GPU does exp / cumulative step / layout (kernel K1z).
    rng="philox" (throughput): counter-based Philox4x32-10 + Box-Muller on the device (kernel K1).  The 64-bit seed
                 is `seed`, or one draw from the global NumPy stream so np.random.seed(...) still makes runs repeatable.
                 store_paths=False (float32 only) keeps NO path matrix: 4 bytes per path instead of 4 (n+1); the backward
                 sweep regenerates the columns from the counters and takes the same decisions as the stored set.
    """
    n_time_steps, n_paths = int(n_time_steps), int(n_paths)
    did = _dtype_id(dtype)
    if not store_paths and (rng != "philox" or did != N.F32):
        raise ValueError("store_paths=False needs rng='philox' and dtype='float32' (the columns are regenerated from "
                         "the Philox counters of the float generator)")
    if rng not in ("numpy", "philox"):
        raise ValueError(f"unknown rng {rng!r}: use 'numpy' or 'philox'")
    ctx = ctx or default_context()
    lo, hi = shard_range(n_paths, ctx.world_size, ctx.rank, 4 if (rng == "philox" and did == N.F32) else 1)
    h = C.c_void_p()
    if not store_paths:
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 62))
        N.check(N.lib().amc_paths_generate_lean(ctx.handle, float(S0), float(r), float(sigma), float(T), n_time_steps,
                                                hi - lo, lo, n_paths, C.c_uint64(int(seed)), C.byref(h)))
        return DevicePaths(ctx, h, hi - lo, n_paths, n_time_steps, did, lo)
    if rng == "numpy":
        if ctx.world_size != 1:
            raise ValueError("rng='numpy' replays the reference's single global stream; use rng='philox' when sharded")
        Z = np.random.normal(size=(n_paths, n_time_steps))                       # amc.py:74
        N.check(N.lib().amc_paths_from_normals(ctx.handle, Z.ctypes.data, float(S0), float(r), float(sigma), float(T),
                                               n_time_steps, n_paths, n_paths, did, C.byref(h)))
        ctx.sync()          # Z must stay alive until the copy has been consumed
    elif rng == "philox":
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 62))
        N.check(N.lib().amc_paths_generate(ctx.handle, float(S0), float(r), float(sigma), float(T), n_time_steps,
                                           hi - lo, lo, n_paths, did, C.c_uint64(int(seed)), C.byref(h)))
    else:
        raise ValueError(f"unknown rng {rng!r}: use 'numpy' or 'philox'")
    return DevicePaths(ctx, h, hi - lo, n_paths, n_time_steps, did, lo)


def paths_from_normals(Z, S0, r, sigma, T, *, dtype="float64", ctx: Context | None = None):
    """Paths from caller-supplied standard normals Z[n_paths, n_time_steps] (host ndarray): the A/B entry."""
    ctx = ctx or default_context()
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    P, n = Z.shape
    h = C.c_void_p()
    did = _dtype_id(dtype)
    N.check(N.lib().amc_paths_from_normals(ctx.handle, Z.ctypes.data, float(S0), float(r), float(sigma), float(T),
                                           n, P, P, did, C.byref(h)))
    ctx.sync()
    return DevicePaths(ctx, h, P, P, n, did)


def paths_from_host(paths, *, dtype="float64", n_paths_global=None, ctx: Context | None = None):
    """Adopt a reference-layout [n_paths, n_time_steps+1] ndarray (e.g. produced by the reference itself)."""
    ctx = ctx or default_context()
    A = np.ascontiguousarray(paths, dtype=np.float64)
    if A.ndim != 2 or A.shape[1] < 1:
        raise ValueError(f"paths must be [n_paths, n_time_steps+1], got shape {A.shape}")
    P, n1 = A.shape
    h = C.c_void_p()
    did = _dtype_id(dtype)
    N.check(N.lib().amc_paths_from_host(ctx.handle, A.ctypes.data, n1 - 1, P, int(n_paths_global or P), did,
                                        C.byref(h)))
    ctx.sync()
    return DevicePaths(ctx, h, P, int(n_paths_global or P), n1 - 1, did)


# --------------------------------------------------------------------------------------------- pricing
class LsmResult:
    """Everything one backward sweep produced."""

    def __init__(self, price, n_time_steps, degree, steps, timing, exercise_steps, cashflow0):
        self.price = price
        self.n_time_steps, self.degree = n_time_steps, degree
        self.gamma, self.beta, self.sv = steps["gamma"], steps["beta"], steps["sv"]
        self.mean_x, self.std_x, self.rank = steps["mean_x"], steps["std_x"], steps["rank"]
        self.pivot_loss = steps["pivot_loss"]
        self.timing = timing
        self.exercise_steps, self.cashflow0 = exercise_steps, cashflow0


def _check_basis(basis_type):
    if basis_type not in _BASES and basis_type not in _EXTRA_BASES:
        # message of amc.py:104
        raise ValueError(f"Unknown basis type '{basis_type}'. Use 'Power', 'Chebyshev', or 'Legendre'.")


def _as_device_paths(paths, ctx):
    if isinstance(paths, DevicePaths):
        return paths, False
    return paths_from_host(paths, ctx=ctx), True


def lsm_price(paths, K, r, dt, option_type, barrier_level=None, exercise_type="European", basis_type="Chebyshev",
              degree=4, scaling=False, scaling_factor=2, *, want_regression=None, want_exercise_steps=False,
              want_cashflows=False, want_svd=False, state_dtype="float64", profile=False,
              ctx: Context | None = None) -> LsmResult:
    """One backward sweep (amc.py:139-197) with all diagnostics.  `lsmc_option_pricing` is the reference-shaped wrapper."""
    ctx = ctx or default_context()
    dp, temporary = _as_device_paths(paths, ctx)
    try:
        n = dp.n_time_steps
        if n >= 1:
            _check_basis(basis_type)          # the reference only reaches amc.py:103 when a regression runs
        american = exercise_type == "American"                                   # amc.py:154
        if want_regression is None:
            want_regression = False
        spec = N.LsmSpec(K=float(K), r=float(r), dt=float(dt),
                         barrier=float("nan") if barrier_level is None else float(barrier_level),   # amc.py:172
                         scaling_factor=float(scaling_factor), is_put=int(option_type == "Put"),    # amc.py:86
                         is_american=int(american), basis=N.BASIS_ID.get(basis_type, 0), degree=int(degree),
                         scaling=int(bool(scaling)), want_regression=int(bool(want_regression)),
                         want_exercise_steps=int(bool(want_exercise_steps)), want_svd=int(bool(want_svd)),
                         state_f32=int(_dtype_id(state_dtype) == N.F32))
        rows = n + 1
        st = dict(gamma=np.zeros((rows, N.AMC_MAX_K)), beta=np.zeros((rows, N.AMC_MAX_K)),
                  sv=np.zeros((rows, N.AMC_MAX_K)), mean_x=np.zeros(rows), std_x=np.zeros(rows),
                  rank=np.zeros(rows, dtype=np.int32), pivot_loss=np.zeros(rows))
        steps = N.LsmSteps(gamma=st["gamma"].ctypes.data_as(N.c_double_p), beta=st["beta"].ctypes.data_as(N.c_double_p),
                           sv=st["sv"].ctypes.data_as(N.c_double_p), mean_x=st["mean_x"].ctypes.data_as(N.c_double_p),
                           std_x=st["std_x"].ctypes.data_as(N.c_double_p), rank=st["rank"].ctypes.data_as(N.c_int_p),
                           pivot_loss=st["pivot_loss"].ctypes.data_as(N.c_double_p))
        timing = N.LsmTiming()
        price = C.c_double()
        ex = np.empty(dp.n_paths_local, dtype=np.int32) if want_exercise_steps else None
        cf = np.empty(dp.n_paths_local, dtype=np.float64) if want_cashflows else None
        N.check(N.lib().amc_lsm_price(ctx.handle, dp.handle, C.byref(spec), C.byref(price), C.byref(steps),
                                      ex.ctypes.data if ex is not None else None,
                                      cf.ctypes.data if cf is not None else None, C.byref(timing), int(bool(profile))))
        tm = dict(total_ms=timing.total_ms, step_kernel_ms=timing.step_kernel_ms, solve_kernel_ms=timing.solve_kernel_ms,
                  step_launches=timing.step_launches, solve_launches=timing.solve_launches,
                  other_launches=timing.other_launches, sweep_kind=timing.sweep_kind)
        res = LsmResult(np.float64(price.value), n, int(degree), st, tm, ex, cf)
        worst = float(st["pivot_loss"].max()) if rows else 0.0
        if worst > PIVOT_LOSS_WARN:
            import warnings
            t_bad = int(st["pivot_loss"].argmax())
            warnings.warn(f"regression at step {t_bad}: the moment-based solve lost {math.log10(worst):.0f} digits "
                          f"factorising the degree-{int(degree)} Gram of a very heavy-tailed column; rank decision and "
                          "fit may deviate from numpy.linalg.lstsq -- lower the degree", RuntimeWarning, stacklevel=3)
        return res
    finally:
        if temporary:
            dp.free()


def lsm_price_batch(paths, contracts, r, dt, barrier_level=None, basis_type="Chebyshev", degree=4, scaling=False,
                    scaling_factor=2, *, want_gamma=False, state_dtype="float64", profile=False,
                    ctx: Context | None = None):
    """Price several contracts on ONE path set in the same launches (include/amc.h: amc_lsm_price_batch).

    `contracts` is a sequence of (K, option_type, exercise_type) -- the arguments of amc.py:180 that may vary inside a
    batch; r, dt, barrier, basis and scaling are shared.  Returns an ndarray of prices (and, with want_gamma, the
    per-contract continuation polynomials [n_contracts, n_time_steps+1, AMC_MAX_K]).  Each price equals what
    `lsmc_option_pricing` returns for that contract alone, to rounding.  The reference has no batched call: its sweeps
    (plots.py:100-107) loop over contracts in Python, re-simulating and re-pricing each one.
    """
    ctx = ctx or default_context()
    dp, temporary = _as_device_paths(paths, ctx)
    try:
        contracts = list(contracts)
        if not contracts:
            return np.zeros(0)
        if dp.n_time_steps >= 1:
            _check_basis(basis_type)
        specs = (N.LsmSpec * len(contracts))()
        for i, (K, option_type, exercise_type) in enumerate(contracts):
            specs[i] = N.LsmSpec(K=float(K), r=float(r), dt=float(dt),
                                 barrier=float("nan") if barrier_level is None else float(barrier_level),
                                 scaling_factor=float(scaling_factor), is_put=int(option_type == "Put"),
                                 is_american=int(exercise_type == "American"), basis=N.BASIS_ID.get(basis_type, 0),
                                 degree=int(degree), scaling=int(bool(scaling)), want_regression=0,
                                 want_exercise_steps=0, want_svd=0, state_f32=int(_dtype_id(state_dtype) == N.F32))
        prices = np.zeros(len(contracts))
        gamma = np.zeros((len(contracts), dp.n_time_steps + 1, N.AMC_MAX_K)) if want_gamma else None
        timing = N.LsmTiming()
        N.check(N.lib().amc_lsm_price_batch(ctx.handle, dp.handle, specs, len(contracts),
                                            prices.ctypes.data_as(N.c_double_p),
                                            gamma.ctypes.data if gamma is not None else None, C.byref(timing),
                                            int(bool(profile))))
        prices_timing = dict(total_ms=timing.total_ms, step_kernel_ms=timing.step_kernel_ms,
                             solve_kernel_ms=timing.solve_kernel_ms, step_launches=timing.step_launches,
                             solve_launches=timing.solve_launches, other_launches=timing.other_launches,
                             sweep_kind=timing.sweep_kind)
        lsm_price_batch.last_timing = prices_timing
        return (prices, gamma) if want_gamma else prices
    finally:
        if temporary:
            dp.free()


class ContinuationValues(Sequence):
    """Lazy stand-in for the list built at amc.py:164-167: item t is (t, paths[:, t], continuation_estimated_t).

    The sweep keeps only the continuation polynomial of every step (k doubles); a step's [n_paths] vectors are
    computed on the device when the item is read.  For European exercise the regressions cannot change the price,
    so they are only run (once) when the sequence is first indexed.
    """

    def __init__(self, paths: DevicePaths, owns_paths, args, result: LsmResult | None, ctx: Context):
        self._paths, self._owns, self._args, self._result, self._ctx = paths, owns_paths, args, result, ctx

    def __len__(self):
        return self._paths.n_time_steps + 1

    def _ensure(self):
        if self._result is None:
            self._result = lsm_price(self._paths, *self._args[0], **self._args[1], want_regression=True, ctx=self._ctx)
        return self._result

    def __getitem__(self, t):
        if isinstance(t, slice):
            return [self[i] for i in range(*t.indices(len(self)))]
        t = int(t)
        if t < 0:
            t += len(self)
        if not 0 <= t < len(self):
            raise IndexError(t)
        n = self._paths.n_time_steps
        S_t = self._paths.column(t)
        if t == n:                                           # amc.py:145: zeros at maturity
            return (t, S_t, np.zeros(self._paths.n_paths_local))
        res = self._ensure()
        out = np.empty(self._paths.n_paths_local)
        gam = np.ascontiguousarray(res.gamma[t])
        N.check(N.lib().amc_continuation(self._ctx.handle, self._paths.handle, t, gam.ctypes.data, res.degree,
                                         out.ctypes.data))
        return (t, S_t, out)

    def __del__(self):
        if getattr(self, "_owns", False):
            try:
                self._paths.free()
            except Exception:
                pass


def lsmc_option_pricing(paths, K, r, dt, option_type, barrier_level=None,
                        exercise_type="European", basis_type="Chebyshev", degree=4,
                        **kwargs):
    """Drop-in for amc.py:180-197: returns (option_price, continuation_values)."""
    extra = set(kwargs) - {"scaling", "scaling_factor"}
    if extra:                                                # what regression_estimate (amc.py:110) would say
        raise TypeError(f"regression_estimate() got an unexpected keyword argument '{sorted(extra)[0]}'")
    ctx = default_context()
    dp, temporary = _as_device_paths(paths, ctx)
    args = ((K, r, dt, option_type, barrier_level, exercise_type, basis_type, degree), dict(kwargs))
    res = lsm_price(dp, *args[0], **args[1], ctx=ctx)
    american = exercise_type == "American"
    cont = ContinuationValues(dp, temporary, args, res if (american or dp.n_time_steps == 0) else None, ctx)
    return res.price, cont


def compute_ccr_exposures(continuation_values):
    """amc.py:400-414: [(t, PFE_5, PFE_95, EPE)] -- per step the 5th / 95th percentile (numpy's linear rule) and the
    mean of the finite continuation values.

    Given the lazy `ContinuationValues` of `lsmc_option_pricing` everything stays on the device: the values are
    recomputed from the stored polynomials and the percentiles found by a radix select (no [n_paths] vector is
    materialised or sorted).  Any other sequence of (t, S_t, values_t) tuples (amc.py:478 passes QuantLib values)
    goes through the same select, one host array at a time.
    """
    if isinstance(continuation_values, ContinuationValues):
        cv = continuation_values
        res = cv._ensure()
        n = cv._paths.n_time_steps
        lo, hi, epe = np.empty(n + 1), np.empty(n + 1), np.empty(n + 1)
        gam = np.ascontiguousarray(res.gamma)
        N.check(N.lib().amc_ccr_exposures(cv._ctx.handle, cv._paths.handle, gam.ctypes.data, res.degree, 0.05, 0.95,
                                          lo.ctypes.data, hi.ctypes.data, epe.ctypes.data))
        return [(t, lo[t], hi[t], epe[t]) for t in range(n + 1)]
    out = []
    ctx = default_context()
    for t, _, cont in continuation_values:
        vals = np.ascontiguousarray(cont, dtype=np.float64).ravel()
        r3 = np.empty(3)
        N.check(N.lib().amc_percentiles(ctx.handle, vals.ctypes.data, vals.size, 0.05, 0.95, r3.ctypes.data))
        out.append((t, r3[0], r3[1], r3[2]))
    return out


# --------------------------------------------------------------------------------------------- small array ops
def intrinsic_value(S, K, option_type="Call"):
    """amc.py:85-86, elementwise on the device; returns float64 with the shape of S."""
    A = np.ascontiguousarray(S, dtype=np.float64)
    out = np.empty_like(A)
    N.check(N.lib().amc_intrinsic_value(default_context().handle, A.ctypes.data, A.size, float(K),
                                        int(option_type == "Put"), out.ctypes.data))
    return out if out.ndim else out[()]


def get_basis_polynomials(X, basis_type, degree):
    """amc.py:98-106: [len(X), degree+1] design matrix."""
    _check_basis(basis_type)
    A = np.ascontiguousarray(X, dtype=np.float64).ravel()
    out = np.empty((A.size, int(degree) + 1))
    N.check(N.lib().amc_basis_matrix(default_context().handle, A.ctypes.data, A.size, N.BASIS_ID[basis_type],
                                     int(degree), out.ctypes.data))
    return out


def regression_estimate(X, Y, basis_type="Power", degree=3, scaling=False, scaling_factor=2):
    """amc.py:110-122: fitted values of the (rank-truncated, numpy-lstsq-equivalent) least-squares fit."""
    _check_basis(basis_type)
    Xa = np.ascontiguousarray(X, dtype=np.float64).ravel()
    Ya = np.ascontiguousarray(Y, dtype=np.float64).ravel()
    if Xa.size != Ya.size:
        raise ValueError(f"X and Y differ in length: {Xa.size} vs {Ya.size}")
    out = np.empty(Xa.size)
    N.check(N.lib().amc_regression_fit(default_context().handle, Xa.ctypes.data, Ya.ctypes.data, Xa.size,
                                       N.BASIS_ID[basis_type], int(degree), int(bool(scaling)), float(scaling_factor),
                                       out.ctypes.data, None, None))
    return out


def apply_exercise(cashflows, exercise_times, in_the_money_idx, exercise_value, continuation_estimated, t):
    """amc.py:90-94 (same positional order): where exercise_value > continuation_estimated (strict) set cashflows /
    exercise_times at `in_the_money_idx`.  Mutates `cashflows` (float64) and `exercise_times` (int64) in place, as the
    reference does."""
    indices = in_the_money_idx
    if not (isinstance(cashflows, np.ndarray) and cashflows.dtype == np.float64 and cashflows.flags.c_contiguous and
            isinstance(exercise_times, np.ndarray) and exercise_times.dtype == np.int64 and exercise_times.flags.c_contiguous):
        raise TypeError("cashflows must be a contiguous float64 ndarray and exercise_times a contiguous int64 ndarray "
                        "(what lsmc_option_pricing allocates, amc.py:185-186)")
    ev = np.ascontiguousarray(exercise_value, dtype=np.float64).ravel()
    ce = np.ascontiguousarray(continuation_estimated, dtype=np.float64).ravel()
    idx = np.ascontiguousarray(indices, dtype=np.int64).ravel()
    if not (ev.size == ce.size == idx.size):
        raise ValueError(f"shape mismatch: {ev.size} exercise values, {ce.size} continuation values, {idx.size} indices")
    try:
        N.check(N.lib().amc_apply_exercise(default_context().handle, cashflows.ctypes.data, exercise_times.ctypes.data,
                                           cashflows.size, ev.ctypes.data, ce.ctypes.data, idx.ctypes.data, idx.size, int(t)))
    except ValueError as e:                       # numpy raises IndexError for a bad fancy index
        raise IndexError(str(e)) from None


def estimate_continuation_values(paths, t, r, dt, cashflows, exercise_times, basis_type, degree, **kwargs):
    """amc.py:126-135: regress the discounted cashflows of ALL paths on paths[:, t]; fitted values clamped at zero."""
    extra = set(kwargs) - {"scaling", "scaling_factor"}
    if extra:
        raise TypeError(f"regression_estimate() got an unexpected keyword argument '{sorted(extra)[0]}'")
    _check_basis(basis_type)
    X = paths.column(t) if isinstance(paths, DevicePaths) else np.ascontiguousarray(np.asarray(paths)[:, t], dtype=np.float64)
    cf = np.ascontiguousarray(cashflows, dtype=np.float64).ravel()
    tau = np.ascontiguousarray(exercise_times, dtype=np.int64).ravel()
    if not (X.size == cf.size == tau.size):
        raise ValueError(f"shape mismatch: {X.size} paths, {cf.size} cashflows, {tau.size} exercise times")
    out = np.empty(X.size)
    N.check(N.lib().amc_estimate_continuation(default_context().handle, X.ctypes.data, cf.ctypes.data, tau.ctypes.data,
                                              X.size, int(t), float(r), float(dt), N.BASIS_ID[basis_type], int(degree),
                                              int(bool(kwargs.get("scaling", False))),
                                              float(kwargs.get("scaling_factor", 2)), out.ctypes.data))
    return out


def perform_backward_iteration(K, r, dt, n_time_steps, barrier_hit, cashflows, paths, option_type, exercise_times,
                               exercise_type, continuation_values, basis_type, degree, **kwargs):
    """amc.py:139-167 with the reference's positional order and in-place contract: fills `cashflows` (undiscounted, float64) and
    `exercise_times` (int64), appends one (t, paths[:, t], continuation_t) tuple per step to `continuation_values` and
    reverses that list (amc.py:164,167).  The whole loop runs as one device sweep.

    Assumes what lsmc_option_pricing passes in (amc.py:185-190): `cashflows` zeros, `exercise_times` full of n, and
    `barrier_hit` the running-OR matrix of precompute_barrier_hit_matrix (once True, True for all later steps); the
    latter is checked.  Materialising every step's vectors is O(n_paths x n_steps) host memory, exactly like the
    reference -- `lsmc_option_pricing` avoids it with a lazy sequence.
    """
    extra = set(kwargs) - {"scaling", "scaling_factor"}
    if extra:
        raise TypeError(f"regression_estimate() got an unexpected keyword argument '{sorted(extra)[0]}'")
    ctx = default_context()
    dp, temporary = _as_device_paths(paths, ctx)
    try:
        n, P = dp.n_time_steps, dp.n_paths_local
        if int(n_time_steps) != n:
            raise ValueError(f"n_time_steps={n_time_steps} but paths has {n + 1} columns")
        hit = np.asarray(barrier_hit, dtype=bool)
        if hit.shape != (P, n + 1):
            raise ValueError(f"barrier_hit must have shape {(P, n + 1)}, got {hit.shape}")
        if n >= 1 and not np.all(hit[:, 1:] >= hit[:, :-1]):
            raise NotImplementedError("barrier_hit must be a running OR over time (precompute_barrier_hit_matrix)")
        first = np.where(hit.any(axis=1), hit.argmax(axis=1), n + 1).astype(np.int32)
        if n >= 1:
            _check_basis(basis_type)
        spec = N.LsmSpec(K=float(K), r=float(r), dt=float(dt), barrier=float("nan"),
                         scaling_factor=float(kwargs.get("scaling_factor", 2)), is_put=int(option_type == "Put"),
                         is_american=int(exercise_type == "American"), basis=N.BASIS_ID.get(basis_type, 0),
                         degree=int(degree), scaling=int(bool(kwargs.get("scaling", False))), want_regression=1,
                         want_exercise_steps=1, want_svd=0, state_f32=0)
        rows = n + 1
        gamma = np.zeros((rows, N.AMC_MAX_K))
        steps = N.LsmSteps(gamma=gamma.ctypes.data_as(N.c_double_p), beta=None, sv=None, mean_x=None, std_x=None, rank=None,
                           pivot_loss=None)
        price = C.c_double()
        tau32 = np.empty(P, dtype=np.int32)
        N.check(N.lib().amc_lsm_price_with_hits(ctx.handle, dp.handle, C.byref(spec), first.ctypes.data, C.byref(price),
                                                C.byref(steps), tau32.ctypes.data, None, None, 0))
        # undiscounted cashflow = payoff at the path's own exercise step where it had knocked in by then (amc.py:93,148)
        s_tau = np.empty(P)
        N.check(N.lib().amc_paths_gather_steps(dp.handle, tau32.ctypes.data, s_tau.ctypes.data))
        pay = intrinsic_value(s_tau, K, option_type)
        cashflows[...] = np.where(first <= tau32, pay, 0.0)
        exercise_times[...] = tau32
        for t in reversed(range(n + 1)):                                               # amc.py:141,164
            S_t = dp.column(t)
            if t == n:
                cont = np.zeros(P)                                                     # amc.py:145
            else:
                cont = np.empty(P)
                g = np.ascontiguousarray(gamma[t])
                N.check(N.lib().amc_continuation(ctx.handle, dp.handle, t, g.ctypes.data, int(degree), cont.ctypes.data))
            continuation_values.append((t, S_t, cont))
        continuation_values.reverse()                                                  # amc.py:167
    finally:
        if temporary:
            dp.free()


def main(params):
    """amc.py:443-503, the computational part: paths -> LSMC price and exposures -> benchmark price -> the same three
    printed lines.  The per-grid-point QuantLib values (amc.py:474, O(n_paths x n_steps) QuantLib calls) and the plots
    are presentation and out of scope; returns what was computed."""
    from .benchmarks import get_quantlib_option
    S0, K, T, r, sigma = params["S0"], params["K"], params["T"], params["r"], params["sigma"]
    n_time_steps, n_paths = params["n_time_steps"], params["n_paths"]
    option_type, exercise_type, barrier_level = params["option_type"], params["exercise_type"], params["barrier_level"]
    paths = generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths)                              # amc.py:465
    dt = T / n_time_steps
    lsmc_price, continuation_values = lsmc_option_pricing(paths, K, r, dt, option_type, barrier_level, exercise_type,
                                                          params["basis_type"], params["degree"],
                                                          scaling=params["scaling"],
                                                          scaling_factor=params["scaling_factor"])   # amc.py:469-471
    lsmc_ccr_exposures = compute_ccr_exposures(continuation_values)                                   # amc.py:479
    benchmark = get_quantlib_option(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type, barrier_level)
    option_description = f"{exercise_type} {option_type}"
    barrier_text = f"with Barrier at {barrier_level}" if barrier_level else "without Barrier"
    print(f"{option_description} Option Price {barrier_text} (LSMC): {lsmc_price:.4f}")               # amc.py:499
    print(f"{option_description} Option Price {barrier_text} (QuantLib): {benchmark.NPV():.4f}")
    out = dict(lsmc_price=float(lsmc_price), benchmark_price=float(benchmark.NPV()), lsmc_ccr_exposures=lsmc_ccr_exposures,
               continuation_values=continuation_values, paths=paths)
    if barrier_level:
        plain = get_quantlib_option(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type)
        print(f"{option_description} Option Price without Barrier (QuantLib): {plain.NPV():.4f}")
        out["benchmark_price_without_barrier"] = float(plain.NPV())
    return out


def precompute_barrier_hit_matrix(paths, barrier_level):
    """amc.py:171-176: bool [n_paths, n_time_steps+1]."""
    ctx = default_context()
    dp, temporary = _as_device_paths(paths, ctx)
    try:
        out = np.empty(dp.shape, dtype=np.uint8)
        b = float("nan") if barrier_level is None else float(barrier_level)
        N.check(N.lib().amc_barrier_hit_matrix(ctx.handle, dp.handle, b, out.ctypes.data))
        return out.view(np.bool_)
    finally:
        if temporary:
            dp.free()
