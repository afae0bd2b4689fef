"""CPU tier, world_size 2 over gloo: the host-side logic of the multi-GPU path.

What runs on the GPUs (per-step all-reduce of the moment sums, identical solve on every rank, local decisions) is
re-enacted with the NumPy emulation of the kernels and a gloo all-reduce, and must equal the single-process result;
the NCCL-id rendezvous helper is exercised with a stand-in id factory.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pipeline_emulator as emu
    from american_monte_carlo_b200 import shard_range
    from american_monte_carlo_b200.distributed import exchange_unique_id
    from oracle import lsm_oracle as orc

    uid = exchange_unique_id(lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))

    S0, K, r, sigma, T, n, P, d = 36.0, 40.0, 0.06, 0.2, 1.0, 20, 20001, 3
    np.random.seed(5)
    paths = orc.generate_asset_paths(S0, r, sigma, T, n, P)          # same on every rank (same seed)
    lo, hi = shard_range(P, world, rank)
    S = np.ascontiguousarray(paths[lo:hi].T)
    mu, sg = emu.column_maps(np.ascontiguousarray(paths.T))           # global maps (the GPU path all-gathers them)
    rdt = r * T / n
    U = np.maximum(K - S[n], 0) * np.exp(-rdt * n)
    tau = np.full(hi - lo, n)
    for t in range(n - 1, -1, -1):
        z = (S[t] - mu[t]) * (1.0 / sg[t])
        h, g = emu.moments(z, U, d)
        buf = torch.from_numpy(np.concatenate([h, g]))
        dist.all_reduce(buf)                                          # the one exchange per step
        tot = buf.numpy()
        res = emu.solve(d, "Power", False, 2, P, tot[:2 * d + 1], tot[2 * d + 1:], np.exp(rdt * t), mu[t], sg[t])
        iv = np.maximum(K - S[t], 0)
        take = (iv > 0) & (iv > emu.horner(res["gamma"], z))
        U = np.where(take, iv * np.exp(-rdt * t), U)
        tau = np.where(take, t, tau)
    s = torch.tensor([U.sum()])
    dist.all_reduce(s)
    want = orc.lsm_backward(paths, K, r, T / n, "Put", None, "American", "Power", d, keep_continuation=False)
    flips = torch.tensor([int((tau != want.exercise_times[lo:hi]).sum())])
    dist.all_reduce(flips)
    if rank == 0:
        q.put((float(s.item()) / P, float(want.price), int(flips.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_moment_allreduce_reproduces_the_oracle_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want, flips = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert flips == 0
    assert abs(got - want) <= 1e-12 * want
