#!/usr/bin/env python
"""Benchmark of the Longstaff-Schwartz hot path (path simulation + LSM backward induction) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c1|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pricing of the workload's contract: build the path set from the step's inputs and run the whole
backward sweep.  Default workload = BASELINE.json configs[1] ("c2"): American put S0=36 K=40 r=0.06 sigma=0.2 T=1,
10M paths x 50 steps per GPU, FP64, degree-3 Power basis, the reference's own seed-42 standard normals injected.
  value : path-steps/s, inputs (the normals Z) already resident in HBM when the timed region starts
  e2e   : the same through the public API with Z in pinned HOST memory, the H2D copy and the D2H of the price inside
          the timed region
N > 1: weak scaling -- every rank prices its own 10M-path shard of ONE contract (global regression: one NCCL
all-reduce of the 10 moment sums per time step); value = all ranks' path-steps / max-over-ranks time.
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LS_PUT = dict(S0=36.0, K=40.0, r=0.06, sigma=0.2, T=1.0)
WORKLOADS = {
    # name: paths per GPU, steps, path dtype, rng, basis, degree, kwargs
    "c1": dict(P=100_000, n=50, dtype="float64", rng="normals", basis="Power", degree=3, kw={},
               label="BASELINE configs[0]: 100k paths x 50 steps, FP64, Power-3, injected normals"),
    "c2": dict(P=10_000_000, n=50, dtype="float64", rng="normals", basis="Power", degree=3, kw={},
               label="BASELINE configs[1]: 10M paths x 50 steps per GPU, FP64, Power-3, injected reference normals"),
    "c3": dict(P=100_000_000, n=252, dtype="float32", state="float32", rng="philox", basis="Power", degree=3, kw={},
               strong=True,
               label="BASELINE configs[2]: 100M paths x 252 steps TOTAL, FP32 paths / FP64 sums, Philox, sharded"),
    "c5": dict(P=50_000_000, n=100, dtype="float32", state="float32", rng="philox", basis="Laguerre", degree=8,
               kw=dict(scaling=True, scaling_factor=2), strong=True,
               label="BASELINE configs[4]: 50M paths x 100 steps TOTAL, degree-8 Laguerre (scaled), Philox"),
    "c4": dict(P=1_000_000, n=50, dtype="float32", state="float32", rng="philox", basis="Power", degree=3, kw={},
               strong=True, grid=True,
               label="BASELINE configs[3]: 1024 contracts (16 strikes x 8 vols x 8 maturities), 1M paths x 50 steps "
                     "each, FP32 paths / FP64 sums, Philox; strikes of a (vol, maturity) cell batched on one path set, "
                     "cells sharded over GPUs"),
}
GOLDEN_C2_PRICE = 4.475181386178888      # tests/golden/golden.json, reference run, seed 42


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, name):
        self.marks = getattr(self, "marks", {})
        self.marks[name] = time.time()

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        marks = getattr(self, "marks", {})
        t_lo, t_hi = marks.get("begin", 0.0), marks.get("end", float("inf"))
        inside = [ln for ts, ln in self.lines if t_lo <= ts <= t_hi + 0.15]
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


def cpu_oracle_rate(wl, sample_paths, repeats=1):
    """The oracle (NumPy restatement of the reference, the only CPU implementation in the repo) timed on this host."""
    import numpy as np
    from oracle import lsm_oracle as orc
    try:
        from threadpoolctl import threadpool_info
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        cores = os.cpu_count() or 1
    best = None
    for _ in range(repeats):
        np.random.seed(42)
        t0 = time.perf_counter()
        paths = orc.generate_asset_paths(LS_PUT["S0"], LS_PUT["r"], LS_PUT["sigma"], LS_PUT["T"], wl["n"], sample_paths)
        t1 = time.perf_counter()
        basis = "Legendre" if wl["basis"] == "Laguerre" else wl["basis"]      # the reference has no Laguerre
        res = orc.lsm_backward(paths, LS_PUT["K"], LS_PUT["r"], LS_PUT["T"] / wl["n"], "Put", None, "American",
                               basis, wl["degree"], keep_continuation=True, **wl["kw"])
        t2 = time.perf_counter()
        del paths
        dt = t2 - t0
        if best is None or dt < best[0]:
            best = (dt, t1 - t0, t2 - t1, float(res.price))
    return dict(value=sample_paths * wl["n"] / best[0], unit="path-steps/s", cores=cores, kind="port",
                sample=f"{sample_paths} paths x {wl['n']} steps (same contract, seed 42): generate_asset_paths "
                       f"{best[1]:.2f} s + lsmc_option_pricing {best[2]:.2f} s; NumPy elementwise ops are single-"
                       f"threaded, lstsq uses {cores} OpenBLAS threads",
                price=best[3], seconds=best[0])


def run_reference(args, wl, rank, emit):
    """--impl reference: the reference's own CPU algorithm (oracle port; /root/reference is absent on the box)."""
    if rank != 0:
        return
    sample = min(wl["P"], 400_000 if wl["n"] <= 100 else 100_000)
    for _ in range(args.warmup):
        cpu_oracle_rate(wl, max(sample // 8, 1000))
    times, rate = [], None
    for _ in range(args.steps):
        rate = cpu_oracle_rate(wl, sample)
        times.append(rate["seconds"])
    mean_s = sum(times) / len(times)
    value = sample * wl["n"] / mean_s
    cb = dict(rate, value=value)
    cb.pop("seconds", None)
    emit(json.dumps({
        "impl": "reference", "metric": "LSM path-steps/sec (path simulation + backward induction)", "value": value,
        "unit": "path-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["label"], "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_contract_grid(args, wl, ctx, world, rank, local_rank, dev, stream, emit):
    """Workload c4: the whole 1024-contract grid is one step (64 path sets, 16 strikes batched on each)."""
    import torch
    import torch.distributed as dist
    from american_monte_carlo_b200 import sweeps

    strikes, vols, mats = sweeps.default_contract_grid()
    n, P = wl["n"], wl["P"]
    n_contracts = len(strikes) * len(vols) * len(mats)
    cells = len(vols) * len(mats)
    my_cells = len([i for i in range(cells) if i % world == rank])
    bS = 4 if wl["dtype"] == "float32" else 8
    bU = 4 if wl["state"] == "float32" else 8

    def one_step(profile=False):
        tm = []
        prices = sweeps.contract_grid(LS_PUT["S0"], LS_PUT["r"], strikes, vols, mats, n, P, "Put", "American", None,
                                      wl["basis"], wl["degree"], seed=42, dtype=wl["dtype"], ctx=ctx, combine=True,
                                      on_cell=lambda iv, im, t: tm.append(t), profile=profile, state_dtype=wl["state"])
        return prices, tm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(steps, profile=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = [one_step(profile) for _ in range(steps)]
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
        sampler.mark("begin")
    ms, out = timed(args.steps)
    _, prof = timed(1, profile=True)
    if rank == 0:
        sampler.mark("end")
    clocks = sampler.stop() if rank == 0 else None

    C = len(strikes)
    unit_steps = float(n_contracts) * P * n
    value = unit_steps * args.steps / (ms * 1e-3)
    tms = prof[0][1]
    step_ms = sum(t["step_kernel_ms"] for t in tms)
    sweep_ms = sum(t["total_ms"] for t in out[-1][1])
    launches = sum(t["step_launches"] for t in tms)
    # batched launch: the two columns are read once for the whole batch, the state once per contract
    alg_bytes = my_cells * P * ((2 * bS + bU * C) + (n - 1) * (2 * bS + 2 * bU * C) + (bS + 2 * bU * C))
    achieved = alg_bytes / (step_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    prices = out[-1][0]
    line = {
        "metric": "LSM path-steps/sec (path simulation + backward induction)", "value": value,
        "unit": "path-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64 sums over f32 paths (%s state)" % wl["state"],
        "data": "synthetic: device Philox4x32-10 + Box-Muller, seed 42 + cell index",
        "config": {"workload": "c4", "description": wl["label"], "contracts": n_contracts,
                   "strikes": [float(strikes[0]), float(strikes[-1]), len(strikes)],
                   "vols": [float(vols[0]), float(vols[-1]), len(vols)],
                   "maturities": [float(mats[0]), float(mats[-1]), len(mats)],
                   "paths_per_contract": P, "time_steps": n, "basis": wl["basis"], "degree": wl["degree"],
                   "path_dtype": wl["dtype"], "state_dtype": wl["state"], "rng": "philox",
                   "unit_definition": "contract x path x step",
                   "l2": "per step and cell the batch streams 16 state vectors (128 MB) -- larger than L2; the two "
                         "4 MB path columns are L2-resident by design"},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": n_contracts * 8, "ms_per_step": ms / args.steps,
                "api": "sweeps.contract_grid = amc_paths_generate + amc_lsm_price_batch per (vol, maturity) cell; no host "
                       "input exists for this workload (device Philox), prices are read back per cell"},
        "gpu_launches": int(args.steps * sum(1 + t["step_launches"] + t["solve_launches"] for t in tms)),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "lsm_step_tma_kernel, grid.y = 16 contracts (batched decision + moments)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / max(launches, 1),
                     "launches_per_sweep": launches, "avg_launch_ms": step_ms / max(launches, 1),
                     "how": "CUDA events around every launch on the launching stream, separate profiled pass"},
        "breakdown_ms": {"sweeps_total": sweep_ms, "step_kernels": step_ms,
                         "pathgen_and_host": ms / args.steps - sweep_ms},
        "price_grid_corners": [float(prices[0, 0, 0]), float(prices[-1, 0, 0]), float(prices[0, -1, -1]),
                               float(prices[-1, -1, -1])],
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_oracle_rate(wl, 1_000_000)
        cb.pop("seconds", None)
        cb["sample"] = "ONE contract of the grid (K=40, sigma=0.2, T=1): " + cb["sample"]
        line["cpu_baseline"] = cb
    if rank == 0:
        emit(json.dumps(line))


def claim_stdout():
    """Route everything libraries print to stdout (e.g. NCCL's version banner) to stderr; return a writer for the
    one JSON line the contract allows on stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(text):
        sys.stdout.flush()
        os.write(real, (text + "\n").encode())
    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--paths", type=int, default=None, help="override paths (per GPU; total for c3/c5)")
    ap.add_argument("--state", default=None, choices=["float64", "float32"],
                    help="storage of the per-path state (default: float64 for FP64 workloads, float32 for FP32-path ones)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.paths:
        wl["P"] = args.paths
    wl["state"] = args.state or wl.get("state", "float64")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, claim_stdout())
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 500)] + sys.argv
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    emit = claim_stdout()

    import numpy as np
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    import american_monte_carlo_b200 as amc
    from american_monte_carlo_b200 import _native as N
    import ctypes as C

    stream = torch.cuda.Stream(device=dev)          # not the legacy default stream: the sweep's launch chain is
    torch.cuda.set_stream(stream)                   # replayed as a CUDA graph, which needs a capturable stream
    ctx = amc.Context(local_rank, stream=stream.cuda_stream)
    amc.set_default_context(ctx)
    if world > 1:
        ids = [amc.Context.new_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_comm(world, rank, ids[0])

    if wl.get("grid"):
        run_contract_grid(args, wl, ctx, world, rank, local_rank, dev, stream, emit)
        if world > 1:
            dist.destroy_process_group()
        return

    n = wl["n"]
    strong = wl.get("strong", False)
    if strong:
        P_global = wl["P"]
        lo, hi = amc.shard_range(P_global, world, rank)
    else:
        P_global = wl["P"] * world
        lo, hi = rank * wl["P"], (rank + 1) * wl["P"]
    P_local = hi - lo
    did = N.F64 if wl["dtype"] == "float64" else N.F32
    bS = 8 if did == N.F64 else 4
    bU = 4 if wl["state"] == "float32" else 8
    dt = LS_PUT["T"] / n
    price_args = (LS_PUT["K"], LS_PUT["r"], dt, "Put", None, "American", wl["basis"], wl["degree"])

    # ---- inputs -------------------------------------------------------------------------------------------
    Z_host = Z_dev = None
    if wl["rng"] == "normals":
        if world == 1 and P_local * n <= 600_000_000:
            np.random.seed(42)                                   # the reference's own stream (amc.py:74)
            Z_np = np.random.normal(size=(P_local, n))
            data = "synthetic: np.random.seed(42); np.random.normal(size=(P, n)) -- the reference's own normals"
        else:
            g = torch.Generator(device=dev)
            g.manual_seed(42 + rank)
            Z_np = None
            data = "synthetic: torch.randn float64 per rank (seed 42+rank)"
        Z_host = torch.empty((P_local, n), dtype=torch.float64, pin_memory=True)
        if Z_np is not None:
            Z_host.numpy()[...] = Z_np
            del Z_np
            Z_dev = Z_host.to(dev, non_blocking=False)
        else:
            Z_dev = torch.randn((P_local, n), dtype=torch.float64, device=dev, generator=g)
            Z_host.copy_(Z_dev)
        torch.cuda.synchronize()
    else:
        data = "synthetic: device Philox4x32-10 + Box-Muller, seed 42 (no host input)"

    lib = N.lib()

    def make_paths(from_host):
        h = C.c_void_p()
        if wl["rng"] == "normals":
            if from_host:
                N.check(lib.amc_paths_from_normals(ctx.handle, Z_host.data_ptr(), LS_PUT["S0"], LS_PUT["r"],
                                                   LS_PUT["sigma"], LS_PUT["T"], n, P_local, P_global, did, C.byref(h)))
            else:
                N.check(lib.amc_paths_from_normals_dev(ctx.handle, Z_dev.data_ptr(), LS_PUT["S0"], LS_PUT["r"],
                                                       LS_PUT["sigma"], LS_PUT["T"], n, P_local, P_global, did,
                                                       C.byref(h)))
        else:
            N.check(lib.amc_paths_generate(ctx.handle, LS_PUT["S0"], LS_PUT["r"], LS_PUT["sigma"], LS_PUT["T"], n,
                                           P_local, lo, P_global, did, C.c_uint64(42), C.byref(h)))
        return amc.DevicePaths(ctx, h, P_local, P_global, n, did, lo)

    def one_step(from_host, profile=False):
        dp = make_paths(from_host)
        res = amc.lsm_price(dp, *price_args, **wl["kw"], state_dtype=wl["state"], profile=profile, ctx=ctx)
        dp.free()
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps, profile=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = []
        for _ in range(steps):
            out.append(one_step(from_host, profile))
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # ---- warm-up, then the timed regions --------------------------------------------------------------------
    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):               # every rank (collectives inside); nvidia-smi needs a moment
        one_step(False)
    if rank == 0:
        time.sleep(0.25)
        sampler.mark("begin")
    ms_dev, res_dev = timed(False, args.steps)
    ms_e2e, res_e2e = timed(True, args.steps)
    # separate pass with CUDA events around every launch of the dominant kernel (not part of `value`)
    _, res_prof = timed(False, max(1, min(args.steps, 3)), profile=True)
    if rank == 0:
        sampler.mark("end")
    clocks = sampler.stop() if rank == 0 else None

    path_steps = float(P_global) * n
    value = path_steps * args.steps / (ms_dev * 1e-3)
    e2e_value = path_steps * args.steps / (ms_e2e * 1e-3)
    tm = res_prof[-1].timing
    step_launches = tm["step_launches"]
    # algorithmic bytes of the fused decide+moments launches of one sweep (DESIGN.md "Kernels"):
    #   maturity launch: read S_n, S_{n-1}, write state                      2 b_S + b_U
    #   n-1 middle launches: read S_t, S_{t-1}, read+write state             2 b_S + 2 b_U
    #   last launch (t = 0): read S_0, read+write state                      b_S + 2 b_U      (b_U = 8 or 4)
    alg_bytes = P_local * ((2 * bS + bU) + (n - 1) * (2 * bS + 2 * bU) + (bS + 2 * bU))
    step_ms = statistics.mean(r.timing["step_kernel_ms"] for r in res_prof)
    solve_ms = statistics.mean(r.timing["solve_kernel_ms"] for r in res_prof)
    sweep_ms = statistics.mean(r.timing["total_ms"] for r in res_dev)
    achieved = alg_bytes / (step_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            tr = json.load(f)
            if tr.get("workload") == args.workload:
                traffic = tr.get("dram_bytes_per_launch")
    except Exception:
        pass

    price = float(res_dev[-1].price)
    line = {
        "metric": "LSM path-steps/sec (path simulation + backward induction)",
        "value": value, "unit": "path-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64" if did == N.F64 else "f64 sums over f32 paths (%s state)" % wl["state"], "data": data,
        "config": {"workload": args.workload, "description": wl["label"], "contract": "American put " + json.dumps(LS_PUT),
                   "paths_per_gpu": P_local, "paths_total": P_global, "time_steps": n, "basis": wl["basis"],
                   "degree": wl["degree"], "path_dtype": wl["dtype"], "state_dtype": wl["state"], "rng": wl["rng"],
                   "allreduce": ctx.transport,
                   "l2": "inputs larger than L2 (path matrix %.1f GB per GPU re-streamed every step)" % (P_local * (n + 1) * bS / 1e9)},
        "e2e": {"value": e2e_value, "unit": "path-steps/s",
                "h2d_bytes_per_step": (P_local * n * 8 if wl["rng"] == "normals" else 0) * world,
                "d2h_bytes_per_step": (8 + (n + 1) * (3 * 11 + 2) * 8 + (n + 1) * 4) * world,
                "ms_per_step": ms_e2e / args.steps,
                "api": "amc_paths_from_normals (pinned host Z) + amc_lsm_price" if wl["rng"] == "normals"
                       else "amc_paths_generate + amc_lsm_price"},
        "gpu_launches": int(args.steps * (1 + tm["step_launches"] + tm["solve_launches"])),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "lsm_step_tma_kernel (fused exercise decision + regression moments)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / step_launches,
                     "launches_per_sweep": step_launches, "avg_launch_ms": step_ms / step_launches,
                     "how": "CUDA events around every launch on the launching stream, separate profiled pass"},
        "breakdown_ms": {"sweep_total": sweep_ms, "step_kernels": step_ms, "solve_kernels": solve_ms,
                         "pathgen": ms_dev / args.steps - sweep_ms},
        "price": price,
    }
    if args.workload == "c2" and world == 1 and wl["P"] == 10_000_000:
        line["price_reference"] = GOLDEN_C2_PRICE
        line["price_rel_err"] = abs(price - GOLDEN_C2_PRICE) / GOLDEN_C2_PRICE
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del Z_host, Z_dev
        sample = 1_000_000 if n <= 60 else 200_000
        cb = cpu_oracle_rate(wl, sample)
        cb.pop("seconds", None)
        line["cpu_baseline"] = cb
    if rank == 0:
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
