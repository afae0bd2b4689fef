"""Sweep drivers: the numeric part of /root/reference/american_monte_carlo_additional_plots.py ("plots.py") and the
strike x vol x maturity contract grid of BASELINE.json config 4, on top of the CUDA hot path.

plots.py runs four sweeps as sequential Python loops of `generate_asset_paths` + `lsmc_option_pricing` and then
plots; the functions here keep the loops' meaning and argument names and return the arrays that were plotted
(plotting itself is out of scope):

    convergence_with_paths        plots.py:22-36     price vs number of paths
    convergence_with_time_steps   plots.py:54-70     price vs number of time steps
    error_heatmap                 plots.py:89-107    |price - benchmark| over (paths x time steps)
    error_vs_basis_degree         plots.py:138-155   price vs basis family and degree 0..max_degree on ONE path set

`contract_grid` is the batched sweep the reference does not have: the strike axis of every (vol, maturity) cell
shares one path set and is priced by one `amc_lsm_price_batch` call; cells are independent, so several GPUs split
the cells (no data-path collective) and the prices are combined by one host all-reduce at the end.
"""
from __future__ import annotations

import numpy as np

from .api import (Context, default_context, generate_asset_paths, lsm_price, lsm_price_batch, lsmc_option_pricing)


def _benchmark(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type, barrier_level):
    from .benchmarks import get_quantlib_option
    return float(get_quantlib_option(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type, barrier_level).NPV())


def convergence_with_paths(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type, barrier_level, path_range,
                           basis_type="Chebyshev", degree=4, **gen_kwargs):
    """plots.py:22-36.  Returns (n_paths_list, lsmc_prices, benchmark_price)."""
    benchmark_price = _benchmark(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type, barrier_level)
    dt = T / n_time_steps
    prices = []
    for n_paths in path_range:
        paths = generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths, **gen_kwargs)
        price, _ = lsmc_option_pricing(paths, K, r, dt, option_type, barrier_level, exercise_type, basis_type, degree)
        prices.append(float(price))
        paths.free()
    return list(path_range), np.array(prices), benchmark_price


def convergence_with_time_steps(S0, K, r, T, sigma, n_paths, option_type, exercise_type, barrier_level,
                                time_step_range, basis_type="Chebyshev", degree=4, **gen_kwargs):
    """plots.py:54-70 (the benchmark uses 10x the finest grid, plots.py:59).  Returns (steps, prices, benchmark)."""
    high_res_steps = max(time_step_range) * 10
    benchmark_price = _benchmark(S0, K, r, T, sigma, high_res_steps, option_type, exercise_type, barrier_level)
    prices = []
    for n_time_steps in time_step_range:
        dt = T / n_time_steps
        paths = generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths, **gen_kwargs)
        price, _ = lsmc_option_pricing(paths, K, r, dt, option_type, barrier_level, exercise_type, basis_type, degree)
        prices.append(float(price))
        paths.free()
    return list(time_step_range), np.array(prices), benchmark_price


def error_heatmap(S0, K, r, T, sigma, time_step_range, path_range, option_type, exercise_type, barrier_level,
                  basis_type="Chebyshev", degree=4, **gen_kwargs):
    """plots.py:89-113.  Returns (error_matrix[len(path_range), len(time_step_range)], (min_n_paths, min_n_time_steps))."""
    high_res_steps = max(time_step_range) * 10
    benchmark_price = _benchmark(S0, K, r, T, sigma, high_res_steps, option_type, exercise_type, barrier_level)
    err = np.zeros((len(path_range), len(time_step_range)))
    for i, n_paths in enumerate(path_range):
        for j, n_time_steps in enumerate(time_step_range):
            dt = T / n_time_steps
            paths = generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths, **gen_kwargs)
            price, _ = lsmc_option_pricing(paths, K, r, dt, option_type, barrier_level, exercise_type,
                                           basis_type=basis_type, degree=degree)
            err[i, j] = abs(float(price) - benchmark_price)
            paths.free()
    i, j = np.unravel_index(np.argmin(err, axis=None), err.shape)
    return err, (path_range[i], time_step_range[j])


def error_vs_basis_degree(S0, K, r, T, sigma, n_time_steps, n_paths, option_type, exercise_type, barrier_level,
                          max_degree, **gen_kwargs):
    """plots.py:138-155: ONE path set, every basis family x degree 0..max_degree.  Returns (degrees, {basis: prices},
    benchmark_price)."""
    benchmark_price = _benchmark(S0, K, r, T, sigma, n_time_steps, option_type, exercise_type, barrier_level)
    paths = generate_asset_paths(S0, r, sigma, T, n_time_steps, n_paths, **gen_kwargs)
    degrees = list(range(0, max_degree + 1))
    out = {}
    for basis_type in ["Chebyshev", "Power", "Legendre"]:
        out[basis_type] = np.array([float(lsm_price(paths, K, r, T / n_time_steps, option_type, barrier_level,
                                                    exercise_type, basis_type, d).price) for d in degrees])
    paths.free()
    return degrees, out, benchmark_price


# ----------------------------------------------------------------------------------------------- config 4
def default_contract_grid():
    """The grid bench.py prices as BASELINE.json config 4 (the reference fixes no such grid):
    16 strikes x 8 vols x 8 maturities = 1024 contracts."""
    strikes = np.linspace(32.0, 48.0, 16)
    vols = np.linspace(0.10, 0.45, 8)
    maturities = np.linspace(0.25, 2.0, 8)
    return strikes, vols, maturities


def cells_of_rank(n_cells, world_size, rank):
    """(vol, maturity) cells priced by `rank`: round-robin, so every rank gets a mix of short and long maturities."""
    return [i for i in range(int(n_cells)) if i % int(world_size) == int(rank)]


def contract_grid(S0, r, strikes, vols, maturities, n_time_steps, n_paths, option_type="Put",
                  exercise_type="American", barrier_level=None, basis_type="Power", degree=3, scaling=False,
                  scaling_factor=2, *, seed=42, dtype="float32", ctx: Context | None = None, combine=True,
                  on_cell=None, profile=False, state_dtype="float32"):
    """Prices[len(strikes), len(vols), len(maturities)] of a strike x vol x maturity grid.

    Paths depend on (sigma, T) only, so every (vol, maturity) cell simulates ONE Philox path set of `n_paths` paths
    (seed + cell index: independent of the number of ranks) and prices all strikes on it in one batched sweep.  Under
    a multi-rank context cell i belongs to rank i % world_size; with combine=True the prices are summed over ranks
    through libamc's communicator so every rank returns the full grid.
    """
    ctx = ctx or default_context()
    strikes = np.asarray(strikes, dtype=float)
    vols = np.asarray(vols, dtype=float)
    maturities = np.asarray(maturities, dtype=float)
    prices = np.zeros((len(strikes), len(vols), len(maturities)))
    world, rank = ctx.world_size, ctx.rank
    mine_cells = set(cells_of_rank(len(vols) * len(maturities), world, rank))
    cell = 0
    for iv, sigma in enumerate(vols):
        for im, T in enumerate(maturities):
            mine = cell in mine_cells
            cell += 1
            if not mine:
                continue
            # one complete path set per cell on this rank (n_paths_global == n_paths_local: contracts are sharded)
            paths = _generate_unsharded(ctx, S0, r, float(sigma), float(T), n_time_steps, n_paths, dtype, seed + cell - 1)
            contracts = [(float(K), option_type, exercise_type) for K in strikes]
            prices[:, iv, im] = lsm_price_batch(paths, contracts, r, float(T) / n_time_steps, barrier_level, basis_type,
                                                degree, scaling, scaling_factor, state_dtype=state_dtype if dtype in ("float32", "f32") else "float64",
                                                profile=profile, ctx=ctx)
            if on_cell is not None:
                on_cell(iv, im, lsm_price_batch.last_timing)
            paths.free()
    if combine and world > 1:
        prices = ctx.allreduce_host(prices.ravel()).reshape(prices.shape)
    return prices


def _generate_unsharded(ctx, S0, r, sigma, T, n_time_steps, n_paths, dtype, seed):
    import ctypes as C

    from . import _native as N
    from .api import DevicePaths, _dtype_id
    h = C.c_void_p()
    did = _dtype_id(dtype)
    N.check(N.lib().amc_paths_generate(ctx.handle, float(S0), float(r), sigma, T, int(n_time_steps), int(n_paths), 0,
                                       int(n_paths), did, C.c_uint64(int(seed)), C.byref(h)))
    return DevicePaths(ctx, h, n_paths, n_paths, n_time_steps, did, 0)
