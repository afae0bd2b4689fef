"""Worker for the multi-GPU parity test: launched by torchrun, one rank per GPU (tests/test_gpu_multi.py)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import american_monte_carlo_b200 as amc  # noqa: E402
from american_monte_carlo_b200.distributed import init_distributed  # noqa: E402
from oracle import lsm_oracle as orc  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=60))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = init_distributed()
    out = {"transport": ctx.transport}

    # (1) injected reference normals, sharded by rows: must equal the oracle with zero flipped decisions
    S0, K, r, sigma, T, n, P = 36.0, 40.0, 0.06, 0.2, 1.0, 50, 100_000
    np.random.seed(42)
    Z = orc.draw_normals(P, n)
    lo, hi = amc.shard_range(P, world, rank)
    from american_monte_carlo_b200 import _native as N
    import ctypes as C
    h = C.c_void_p()
    Zs = np.ascontiguousarray(Z[lo:hi])
    N.check(N.lib().amc_paths_from_normals(ctx.handle, Zs.ctypes.data, S0, r, sigma, T, n, hi - lo, P, N.F64, C.byref(h)))
    ctx.sync()
    dp = amc.DevicePaths(ctx, h, hi - lo, P, n, N.F64, lo)
    res = amc.lsm_price(dp, K, r, T / n, "Put", None, "American", "Power", 3, want_exercise_steps=True, ctx=ctx)
    paths = orc.paths_from_normals(Z, S0, r, sigma, T)
    want = orc.lsm_backward(paths, K, r, T / n, "Put", None, "American", "Power", 3, keep_continuation=False,
                            keep_diag=True)
    flips = torch.tensor([int((res.exercise_steps != want.exercise_times[lo:hi]).sum())], device="cuda")
    dist.all_reduce(flips)
    out["injected"] = dict(price=float(res.price), oracle=float(want.price), flips=int(flips.item()),
                           ranks_equal=res.rank[:n].tolist() == [want.steps[t]["rank"] for t in range(n)])

    # (2) Philox: the union of the shards is the single-GPU path set -> same price up to summation order
    Pp = 1_000_003
    dq = amc.generate_asset_paths(S0, r, sigma, T, n, Pp, rng="philox", seed=11, dtype="float32", ctx=ctx)
    rq = amc.lsm_price(dq, K, r, T / n, "Put", None, "American", "Power", 3, ctx=ctx)
    gam_multi = rq.gamma.copy()
    dq.free()
    if rank == 0:
        solo = amc.Context(local)
        ds = amc.generate_asset_paths(S0, r, sigma, T, n, Pp, rng="philox", seed=11, dtype="float32", ctx=solo)
        rs = amc.lsm_price(ds, K, r, T / n, "Put", None, "American", "Power", 3, ctx=solo)
        out["philox"] = dict(multi=float(rq.price), single=float(rs.price),
                             gamma_max_rel=float(np.max(np.abs(gam_multi - rs.gamma) / (np.abs(rs.gamma) + 1e-300))))
        ds.free()
    # (2b) exposures (amc.py:400-414) of the sharded set: histograms all-reduced -> same percentiles as one GPU
    from american_monte_carlo_b200.api import ContinuationValues
    pargs = ((K, r, T / n, "Put", None, "American", "Power", 3), {})
    dq2 = amc.generate_asset_paths(S0, r, sigma, T, n, Pp, rng="philox", seed=11, dtype="float32", ctx=ctx)
    rq2 = amc.lsm_price(dq2, *pargs[0], want_regression=True, ctx=ctx)
    ex_multi = amc.compute_ccr_exposures(ContinuationValues(dq2, False, pargs, rq2, ctx))
    if rank == 0:
        solo2 = amc.Context(local)
        ds2 = amc.generate_asset_paths(S0, r, sigma, T, n, Pp, rng="philox", seed=11, dtype="float32", ctx=solo2)
        rs2 = amc.lsm_price(ds2, *pargs[0], want_regression=True, ctx=solo2)
        ex_solo = amc.compute_ccr_exposures(ContinuationValues(ds2, False, pargs, rs2, solo2))
        scale = max(abs(e[2]) for e in ex_solo) + 1e-12
        out["exposures"] = dict(
            max_pct_diff=float(max(max(abs(a[1] - b[1]), abs(a[2] - b[2])) for a, b in zip(ex_multi, ex_solo)) / scale),
            max_mean_diff=float(max(abs(a[3] - b[3]) for a, b in zip(ex_multi, ex_solo)) / scale), steps=len(ex_multi))
        ds2.free()
    dq2.free()
    # every rank must hold the same polynomial (the regression is global)
    g = torch.tensor(gam_multi, device="cuda")
    gmax, gmin = g.clone(), g.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
    out["gamma_identical_across_ranks"] = bool(torch.equal(gmax, gmin))

    # (3) adopted host matrix with measured (all-gathered) column maps
    hp = C.c_void_p()
    Ps = np.ascontiguousarray(paths[lo:hi])
    N.check(N.lib().amc_paths_from_host(ctx.handle, Ps.ctypes.data, n, hi - lo, P, N.F64, C.byref(hp)))
    dh = amc.DevicePaths(ctx, hp, hi - lo, P, n, N.F64, lo)
    mu, sg = dh.column_maps()
    rh = amc.lsm_price(dh, K, r, T / n, "Put", None, "American", "Power", 3, ctx=ctx)
    out["adopted"] = dict(price=float(rh.price), mu_err=float(np.max(np.abs(mu - paths.mean(axis=0)) / mu)),
                          sg_err=float(np.max(np.abs(sg[1:] - paths.std(axis=0)[1:]) / sg[1:])))
    # (2c) path-free sets shard like stored ones: the persistent sweep kernel exchanges the sums from inside (peer memory)
    if ctx.transport == "p2p":
        dl = amc.generate_asset_paths(S0, r, sigma, T, n, Pp, rng="philox", seed=11, dtype="float32", store_paths=False, ctx=ctx)
        rl = amc.lsm_price(dl, K, r, T / n, "Put", None, "American", "Power", 3, ctx=ctx)
        dl.free()
        out["lean"] = dict(multi=float(rl.price), stored_multi=float(rq.price))
    # (4) a host ndarray handed to the reference-shaped entry point under a multi-rank default context is a rank-LOCAL
    # path set (n_global == n_local): no exchange, every rank prices all of it by itself and gets the oracle's price
    small = np.ascontiguousarray(paths[:20_000])
    want_small = orc.lsm_backward(small, K, r, T / n, "Put", None, "American", "Power", 3, keep_continuation=False)
    got_small, _ = amc.lsmc_option_pricing(small, K, r, T / n, "Put", None, "American", "Power", 3)
    if rank == 1:        # only one rank calls a second time: a stray exchange would wait for the others and time out
        again, _ = amc.lsmc_option_pricing(small, K, r, T / n, "Put", None, "American", "Power", 3)
        assert again == got_small
    errs = torch.tensor([abs(float(got_small) - float(want_small.price)) / float(want_small.price)], device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    out["local_ndarray"] = dict(max_rel_err=float(errs.item()))
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
