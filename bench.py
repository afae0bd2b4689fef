#!/usr/bin/env python
"""Benchmark of the Longstaff-Schwartz hot path (path simulation + LSM backward induction) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c1|c4|c5] [--lean]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pricing of the workload's contract: build the path set from the step's inputs and run the whole
backward sweep.  Default workload = BASELINE.json configs[1] ("c2"): American put S0=36 K=40 r=0.06 sigma=0.2 T=1,
10M paths x 50 steps per GPU, FP64, degree-3 Power basis, the reference's own seed-42 standard normals injected.
  value : path-steps/s, inputs (the normals Z) already resident in HBM when the timed region starts
  e2e   : the same through the public API with Z in pinned HOST memory, the H2D copy and the D2H of the price inside
          the timed region
N > 1: weak scaling -- every rank prices its own 10M-path shard of ONE contract (global regression: one all-reduce of
the 10 moment sums per time step, fused into the sweep kernel over NVLink peer memory); value = all ranks' path-steps /
max-over-ranks time.

The default run also times BASELINE.json configs[2] -- the north-star configuration, 100M paths x 252 steps IN TOTAL,
FP32 paths / FP64 sums, device Philox, strong scaling over the N GPUs -- and reports it as the block `north_star_c3`
of the same JSON line (its own step count, stated there), including the check that the N-GPU price equals the
committed 1-GPU price of the same seed (`price_matches_n1`).
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import math
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
GOLDEN_C2_PRICE = 4.475181386178888      # tests/golden/golden.json (reference run, seed 42, 10M x 50)
GOLDEN_C3_REDUCED = 4.484746999222944    # tests/golden/golden.json c3_reduced (reference run, seed 42, 1M x 252)
GOLDEN_C5_REDUCED = 4.489059534200499    # tests/golden/golden.json c5_reduced (reference run, 500k x 100, Legendre-8 scaled)
# 1-GPU prices of the Philox workloads with seed 42 and Philox4x32-10 (measured on a B200, profiles/r2_*): the price of
# the same seed on N GPUs must agree to summation-order rounding, because the counters are global path ids
PRICE_N1 = {("c3", 100_000_000, "float32"): None, ("c3", 100_000_000, "float64"): None}
try:
    with open(os.path.join(ROOT, "profiles", "price_n1.json")) as _f:
        for _k, _v in json.load(_f).items():
            _w, _p, _s = _k.split("|")
            PRICE_N1[(_w, int(_p), _s)] = _v
except Exception:
    pass
FP64_FMA_PEAK = 148 * 64 * 1.965e9       # DFMA/s of a B200 at 1965 MHz (64 FP64 lanes per SM)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload):
    """DRAM bytes per sweep from the committed ncu capture of this workload (profiles/summarize_ncu.py writes the file
    from `ncu --set full --cache-control none` reports of >= 10 consecutive launches); None when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            tr = json.load(f)
        rec = tr.get(workload) if isinstance(tr.get(workload), dict) else None
        return rec
    except Exception:
        return None


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


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_rate(wl, sample_paths):
    """The oracle (NumPy restatement of the reference, the only CPU implementation in the repo) timed on this host,
    stage by stage: the draw of the normals (which the GPU `value` arm receives ready-made), the path construction and
    the backward induction."""
    import numpy as np
    from oracle import lsm_oracle as orc
    try:
        from threadpoolctl import threadpool_info
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        cores = os.cpu_count() or 1
    np.random.seed(42)
    t0 = time.perf_counter()
    Z = orc.draw_normals(sample_paths, wl["n"])
    t1 = time.perf_counter()
    paths = orc.paths_from_normals(Z, LS_PUT["S0"], LS_PUT["r"], LS_PUT["sigma"], LS_PUT["T"])
    del Z
    t2 = time.perf_counter()
    basis = "Legendre" if wl["basis"] == "Laguerre" else wl["basis"]      # the reference has no Laguerre
    res = orc.lsm_backward(paths, LS_PUT["K"], LS_PUT["r"], LS_PUT["T"] / wl["n"], "Put", None, "American",
                           basis, wl["degree"], keep_continuation=True, **wl["kw"])
    t3 = time.perf_counter()
    del paths
    work = sample_paths * wl["n"]
    same_work_s = (t2 - t1) + (t3 - t2)           # what the GPU `value` arm does: normals -> paths -> sweep
    return dict(value=work / same_work_s, unit="path-steps/s", cores=cores, kind="port",
                sample=f"{sample_paths} paths x {wl['n']} steps (same contract, seed 42): normals -> paths "
                       f"{t2 - t1:.2f} s + lsmc_option_pricing {t3 - t2:.2f} s (the work of the GPU `value` arm); drawing "
                       f"the normals (np.random.normal) {t1 - t0:.2f} s more; NumPy elementwise ops are single-threaded, "
                       f"lstsq uses {cores} OpenBLAS threads",
                value_including_rng=work / (t3 - t0), draw_normals_s=t1 - t0, paths_s=t2 - t1, lsm_s=t3 - t2,
                price=float(res.price), seconds=same_work_s)


def run_reference(args, wl, rank, emit):
    """--impl reference: the reference's own CPU algorithm (oracle port; /root/reference is absent on the box), on the
    same work the GPU `value` arm does (normals already drawn), at the largest sample that keeps K + W steps within a few
    minutes; the rate at a quarter of that sample is printed beside it so the size trend is visible."""
    if rank != 0:
        return
    if wl["n"] <= 100:
        sample = min(wl["P"], 2_000_000)
    else:
        sample = min(wl["P"], 400_000)
    budget_s = 200.0
    probe = cpu_oracle_rate(wl, max(sample // 4, 1000))
    est = probe["seconds"] * 4.6 + probe["draw_normals_s"] * 4
    while sample > 100_000 and est * args.steps > budget_s:
        sample //= 2
        est /= 2
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(wl, max(sample // 16, 1000))
    times, rate = [], None
    for _ in range(args.steps):
        rate = cpu_oracle_rate(wl, sample)
        times.append(rate["seconds"])
    mean_s = sum(times) / len(times)
    value = sample * wl["n"] / mean_s
    cb = dict(rate, value=value)
    cb.pop("seconds", None)
    cb["rate_at_quarter_sample"] = dict(value=probe["value"], sample=probe["sample"].split(" (")[0])
    emit(json.dumps({
        "impl": "reference", "metric": "LSM path-steps/sec (path simulation + backward induction)", "value": value,
        "unit": "path-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["label"], "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------- GPU arm
class Runtime:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import datetime
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=300))
        import american_monte_carlo_b200 as amc
        self.amc = amc
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.ctx = amc.Context(self.local_rank, stream=self.stream.cuda_stream)
        amc.set_default_context(self.ctx)
        if self.world > 1:
            ids = [amc.Context.new_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            self.ctx.init_comm(self.world, self.rank, ids[0])

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(rt, name, wl, steps, warmup, lean=False, want_e2e=True, sampler=None):
    """Warm up, then time `steps` pricings of workload `wl` (device-resident inputs, then host inputs); returns the
    measurements of the JSON line.  Every number is taken with CUDA events on the launching stream, max over ranks."""
    import ctypes as C
    import numpy as np
    torch, amc = rt.torch, rt.amc
    from american_monte_carlo_b200 import _native as N
    world, rank, ctx = rt.world, rt.rank, rt.ctx
    n = wl["n"]
    strong = wl.get("strong", False)
    unit = 4 if wl["rng"] == "philox" and wl["dtype"] == "float32" else 1
    if strong:
        P_global = wl["P"]
        lo, hi = amc.shard_range(P_global, world, rank, unit)
    else:
        P_global = wl["P"] * world
        lo, hi = rank * wl["P"], (rank + 1) * wl["P"]
    P_local = hi - lo
    did = N.F64 if wl["dtype"] == "float64" else N.F32
    bS = 8 if did == N.F64 else 4
    bU = 4 if wl["state"] == "float32" else 8
    dt = LS_PUT["T"] / n
    price_args = (LS_PUT["K"], LS_PUT["r"], dt, "Put", None, "American", wl["basis"], wl["degree"])

    Z_host = Z_dev = None
    if wl["rng"] == "normals":
        if world == 1 and P_local * n <= 600_000_000:
            np.random.seed(42)                                   # the reference's own stream (amc.py:74)
            Z_np = np.random.normal(size=(P_local, n))
            data = "synthetic: np.random.seed(42); np.random.normal(size=(P, n)) -- the reference's own normals"
        else:
            g = torch.Generator(device=rt.dev)
            g.manual_seed(42 + rank)
            Z_np = None
            data = "synthetic: torch.randn float64 per rank (seed 42+rank)"
        Z_host = torch.empty((P_local, n), dtype=torch.float64, pin_memory=True)
        if Z_np is not None:
            Z_host.numpy()[...] = Z_np
            del Z_np
            Z_dev = Z_host.to(rt.dev, non_blocking=False)
        else:
            Z_dev = torch.randn((P_local, n), dtype=torch.float64, device=rt.dev, generator=g)
            Z_host.copy_(Z_dev)
        torch.cuda.synchronize()
    else:
        data = "synthetic: device Philox4x32-10 + Box-Muller, seed 42, counters = global path ids (no host input)"
    lib = N.lib()

    def make_paths(from_host):
        h = C.c_void_p()
        if wl["rng"] == "normals":
            fn = lib.amc_paths_from_normals if from_host else lib.amc_paths_from_normals_dev
            ptr = Z_host.data_ptr() if from_host else Z_dev.data_ptr()
            N.check(fn(ctx.handle, ptr, LS_PUT["S0"], LS_PUT["r"], LS_PUT["sigma"], LS_PUT["T"], n, P_local, P_global, did,
                       C.byref(h)))
        elif lean:
            N.check(lib.amc_paths_generate_lean(ctx.handle, LS_PUT["S0"], LS_PUT["r"], LS_PUT["sigma"], LS_PUT["T"], n,
                                                P_local, lo, P_global, C.c_uint64(42), C.byref(h)))
        else:
            N.check(lib.amc_paths_generate(ctx.handle, LS_PUT["S0"], LS_PUT["r"], LS_PUT["sigma"], LS_PUT["T"], n,
                                           P_local, lo, P_global, did, C.c_uint64(42), C.byref(h)))
        return amc.DevicePaths(ctx, h, P_local, P_global, n, did, lo)

    def one_step(from_host, profile=False):
        dp = make_paths(from_host)
        res = amc.lsm_price(dp, *price_args, **wl["kw"], state_dtype=wl["state"], profile=profile, ctx=ctx)
        dp.free()
        return res

    def timed(from_host, k, profile=False):
        rt.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(rt.stream)
        out = [one_step(from_host, profile) for _ in range(k)]
        e1.record(rt.stream)
        rt.barrier()
        return rt.max_over_ranks(e0.elapsed_time(e1)), out

    for _ in range(warmup):
        one_step(False)
    if sampler is not None and rank == 0:
        time.sleep(0.25)
        sampler.mark("begin")
    ms_dev, res_dev = timed(False, steps)
    if want_e2e:
        ms_e2e, res_e2e = timed(True, steps)
    else:
        ms_e2e = ms_dev
    if sampler is not None and rank == 0:
        sampler.mark("end")

    res_kernel = res_dev
    if res_dev[-1].timing["step_kernel_ms"] <= 0.0:
        # per-step launch chain (contract batches, NCCL transport, AMC_PERSISTENT=0): the kernel times come from a separate
        # pass with CUDA events around every launch
        _, res_kernel = timed(False, max(1, min(steps, 3)), profile=True)
    if sampler is not None and rank == 0:
        sampler.mark("end")
    path_steps = float(P_global) * n
    value = path_steps * steps / (ms_dev * 1e-3)
    tm = res_kernel[-1].timing
    # algorithmic bytes of the backward sweep (SURVEY.md section 8d / DESIGN.md "Kernels"), per rank:
    #   maturity pass: read S_n, S_{n-1}, write state                 2 b_S + b_U
    #   n-1 middle passes: read S_t, S_{t-1}, read+write state        2 b_S + 2 b_U
    #   last pass (t = 0): read S_0, read+write state                 b_S + 2 b_U
    # path-free sets: the stored column pair is replaced by the log-price state (read + write 4 B per pass)
    if lean:
        alg_bytes = P_local * ((4 + bU) + (n - 1) * (8 + 2 * bU) + (4 + 2 * bU))
    else:
        alg_bytes = P_local * ((2 * bS + bU) + (n - 1) * (2 * bS + 2 * bU) + (bS + 2 * bU))
    step_ms = rt.max_over_ranks(statistics.mean(r.timing["step_kernel_ms"] for r in res_kernel))
    sweep_ms = rt.max_over_ranks(statistics.mean(r.timing["total_ms"] for r in res_dev))
    launches = max(tm["step_launches"], 1)
    achieved = alg_bytes / (step_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    D = wl["degree"]
    fp64_bound = D >= 5
    # FP64 work of the streaming loop: (3d + 2) fused multiply-adds per path-step for the moments and the Horner test
    # plus ~8 for payoff / maps / compares (DESIGN.md "Kernels")
    dfma_per_path_step = 3 * D + 2 + 8 + D
    fma_rate = P_local * float(n) * dfma_per_path_step / (step_ms * 1e-3)
    price = float(res_dev[-1].price)
    kern = {1: "lsm_sweep_kernel (persistent: exercise decision + regression moments of all passes, solve by the last "
               "block of every pass)",
            2: "lsm_cluster_kernel (one 16-CTA cluster: state and columns in shared memory, all passes, solve in every CTA)"
            }.get(tm.get("sweep_kind", 0), "lsm_step_tma_kernel (fused exercise decision + regression moments)")
    out = dict(
        value=value, ms_per_step=ms_dev / steps, steps=steps, warmup=warmup, scaling="strong" if strong else "weak",
        dtype="f64" if did == N.F64 else "f64 sums over f32 paths (%s state)" % wl["state"], data=data,
        config={"workload": name, "description": wl["label"], "contract": "American put " + json.dumps(LS_PUT),
                "paths_per_gpu": P_local, "paths_total": P_global, "time_steps": n, "basis": wl["basis"],
                "degree": wl["degree"], "path_dtype": wl["dtype"], "state_dtype": wl["state"], "rng": wl["rng"],
                "store_paths": not lean, "allreduce": ctx.transport,
                "l2": "inputs larger than L2 (%s %.2f GB per GPU re-streamed every step)" %
                      ("state" if lean else "path matrix", (P_local * (4 + bU) if lean else P_local * (n + 1) * bS) / 1e9)},
        e2e={"value": path_steps * steps / (ms_e2e * 1e-3), "unit": "path-steps/s",
             "h2d_bytes_per_step": (P_local * n * 8 if wl["rng"] == "normals" else 0) * world,
             "d2h_bytes_per_step": (8 + (n + 1) * (3 * 11 + 2) * 8 + (n + 1) * 4) * world,
             "ms_per_step": ms_e2e / steps,
             "api": "amc_paths_from_normals (pinned host Z) + amc_lsm_price" if wl["rng"] == "normals"
                    else "amc_paths_generate%s + amc_lsm_price (no host input exists: device Philox)" % ("_lean" if lean else "")},
        gpu_launches=int(steps * (1 + tm["step_launches"] + tm["solve_launches"] + tm["other_launches"])),
        roofline={"bound": "fp64" if fp64_bound else "hbm", "kernel": kern,
                  "achieved": fma_rate / 1e12 if fp64_bound else achieved,
                  "peak": FP64_FMA_PEAK / 1e12 if fp64_bound else peak,
                  "unit": "TDFMA/s" if fp64_bound else "GB/s",
                  "frac": fma_rate / FP64_FMA_PEAK if fp64_bound else achieved / peak,
                  "traffic": None,
                  "peak_source": "148 SMs x 64 FP64 lanes x 1.965 GHz" if fp64_bound else peak_src,
                  "hbm_achieved_gbs": achieved, "hbm_frac": achieved / peak,
                  "algorithmic_bytes_per_launch": alg_bytes / launches, "launches_per_sweep": launches,
                  "avg_launch_ms": step_ms / launches,
                  "how": "CUDA events around the sweep kernel on the launching stream, inside the timed region "
                         "(mean over its steps, max over ranks)"},
        breakdown_ms={"sweep_total": sweep_ms, "sweep_kernel": step_ms, "pathgen": ms_dev / steps - sweep_ms,
                      "per_time_step_us": 1e3 * step_ms / (n + 1)},
        price=price)
    tr = measured_traffic(name)
    if tr and tr.get("paths_per_gpu") == P_local and not lean:
        # per launch, like `achieved`: the chain launches once per time step, the persistent kernel once per sweep
        out["roofline"]["traffic"] = tr.get("dram_bytes_per_launch") if launches > 1 else tr.get("dram_bytes_per_sweep")
        out["roofline"]["traffic_frac_of_peak"] = (tr.get("dram_bytes_per_launch") / (step_ms / launches * 1e-3) / 1e9 / peak
                                                   if launches > 1 else None)
        out["roofline"]["traffic_source"] = tr.get("source")
    del Z_host, Z_dev
    return out


def price_checks(name, wl, res, world, lean):
    """In-line correctness of the timed result: against the reference's number where the inputs are the reference's,
    within Monte Carlo error of the reference's reduced-size run otherwise, and -- multi-GPU -- against the 1-GPU price."""
    out = {}
    price = res["price"]
    P = res["config"]["paths_total"]
    if name == "c2" and world == 1 and P == 10_000_000:
        out["price_reference"] = GOLDEN_C2_PRICE
        out["price_rel_err"] = abs(price - GOLDEN_C2_PRICE) / GOLDEN_C2_PRICE
    ref = {"c2": (GOLDEN_C2_PRICE, 10_000_000), "c1": (GOLDEN_C2_PRICE, 10_000_000), "c3": (GOLDEN_C3_REDUCED, 1_000_000),
           "c5": (GOLDEN_C5_REDUCED, 500_000)}.get(name)
    if ref and "price_rel_err" not in out:
        # cashflow standard deviation of this put is ~3.1 (measured); both samples carry Monte Carlo error
        se = 3.1 * math.sqrt(1.0 / P + 1.0 / ref[1])
        out["price_check"] = {"independent_sample": ref[0], "its_paths": ref[1], "diff": price - ref[0], "mc_standard_error": se,
                              "within_4_se": abs(price - ref[0]) < 4 * se}
    key = (name, P, wl["state"])
    if wl["rng"] == "philox" and PRICE_N1.get(key) is not None:
        p1 = PRICE_N1[key]
        out["price_n1"] = p1
        out["price_matches_n1"] = abs(price - p1) <= 1e-9 * abs(p1)
    return out


def run_contract_grid(args, wl, rt, emit):
    """Workload c4: the whole 1024-contract grid is one step (64 path sets, 16 strikes batched on each)."""
    torch = rt.torch
    from american_monte_carlo_b200 import sweeps
    ctx, world, rank = rt.ctx, rt.world, rt.rank
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

    def timed(steps, profile=False):
        rt.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(rt.stream)
        out = [one_step(profile) for _ in range(steps)]
        e1.record(rt.stream)
        rt.barrier()
        return rt.max_over_ranks(e0.elapsed_time(e1)), out

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(rt.local_rank)
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
    # FP64 work per contract-path-step at degree 3: ~24 DFMA-class instructions (decision 8 + cross sums 7 + shared powers)
    fma_rate = my_cells * float(C) * P * n * 22.0 / (step_ms * 1e-3)
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
                   "l2": "per step and cell the batch streams 16 state vectors (64-128 MB) -- about the size of L2; the "
                         "two 4 MB path columns are L2-resident by design"},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": n_contracts * 8, "ms_per_step": ms / args.steps,
                "api": "sweeps.contract_grid = amc_paths_generate + amc_lsm_price_batch per (vol, maturity) cell; no host "
                       "input exists for this workload (device Philox), prices are read back per cell"},
        "gpu_launches": int(args.steps * sum(1 + t["step_launches"] + t["solve_launches"] for t in tms)),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "kernel": "lsm_step_tma_kernel, grid.y = 16 contracts (batched decision + moments)",
                     "achieved": fma_rate / 1e12, "peak": FP64_FMA_PEAK / 1e12, "unit": "TDFMA/s",
                     "frac": fma_rate / FP64_FMA_PEAK, "traffic": None,
                     "peak_source": "148 SMs x 64 FP64 lanes x 1.965 GHz; the batch is FP64-pipe bound (columns and most of "
                                    "the state are L2 hits), the byte figures are kept for reference",
                     "hbm_achieved_gbs": achieved, "hbm_frac": achieved / peak, "hbm_peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes / max(launches, 1),
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
    ap.add_argument("--lean", action="store_true", help="Philox workloads: path-free set (store_paths=False)")
    ap.add_argument("--scaling", action="store_true",
                    help="regress on the standardised column (regression_estimate(scaling=True, scaling_factor=2))")
    ap.add_argument("--degree", type=int, default=None, help="override the basis degree of the workload (experiments)")
    ap.add_argument("--basis", default=None, help="override the basis family of the workload (experiments)")
    ap.add_argument("--no-c3", action="store_true", help="default workload only: skip the north_star_c3 block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.paths:
        wl["P"] = args.paths
    wl["state"] = args.state or wl.get("state", "float64")
    if args.scaling:
        wl["kw"] = dict(scaling=True, scaling_factor=2)
    if args.degree is not None:
        wl["degree"] = args.degree
    if args.basis:
        wl["basis"] = args.basis
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
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
    rt = Runtime(args)

    if wl.get("grid"):
        run_contract_grid(args, wl, rt, emit)
        rt.close()
        return
    if args.lean and wl["rng"] != "philox":
        raise SystemExit("--lean needs a Philox workload (c3, c5)")

    sampler = ClockSampler(rt.local_rank)
    if rank == 0:
        sampler.start()
    res = measure(rt, args.workload, wl, args.steps, max(args.warmup, 3), lean=args.lean, sampler=sampler)
    line = {"metric": "LSM path-steps/sec (path simulation + backward induction)", "value": res.pop("value"),
            "unit": "path-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": res.pop("ms_per_step"), "higher_is_better": True, "scaling": res.pop("scaling"),
            "vs_baseline": None}
    res.pop("steps"); res.pop("warmup")
    line.update(res)
    line.update(price_checks(args.workload, wl, line, world, args.lean))

    if args.workload == "c2" and not args.paths and not args.no_c3:
        # the north-star configuration in the same run: 100M paths x 252 steps in total, strong scaling
        try:
            c3 = dict(WORKLOADS["c3"])
            c3["state"] = c3.get("state", "float32")
            k3 = max(3, min(args.steps, 5))
            r3 = measure(rt, "c3", c3, k3, 3, lean=False, want_e2e=False)
            peak, peak_src = measured_peak()
            total_bytes = float(c3["P"]) * c3["n"] * 20.0     # generation 4 B + sweep 16 B per path-step
            r3["end_to_end_hbm"] = {"algorithmic_bytes_per_path_step": 20,
                                    "achieved_gbs": total_bytes / (r3["ms_per_step"] * 1e-3) / 1e9,
                                    "aggregate_peak_gbs": peak * world, "frac_of_aggregate_copy_bandwidth":
                                    total_bytes / (r3["ms_per_step"] * 1e-3) / 1e9 / (peak * world), "peak_source": peak_src}
            r3.update(price_checks("c3", c3, r3, world, False))
            r3["n_gpus"] = world
            r3["unit"] = "path-steps/s"
            r3.pop("e2e", None)
            line["north_star_c3"] = r3
        except Exception as exc:            # the headline line must survive a failure of the second workload
            line["north_star_c3"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
    if rank == 0:
        sampler.mark("end")
        line["clocks"] = sampler.stop()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_oracle_rate(wl, 1_000_000 if wl["n"] <= 60 else 200_000)
        cb.pop("seconds", None)
        line["cpu_baseline"] = cb
    if rank == 0:
        emit(json.dumps(line))
    rt.close()


if __name__ == "__main__":
    main()
