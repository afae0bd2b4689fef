"""GPU tier: parity at the reduced shapes of BASELINE.json configs[2] and configs[4].

    c3_reduced  1M paths x 252 steps, Power-3            reference price 4.4847469992...
    c5_reduced  500k paths x 100 steps, Legendre-8 scaled (sf=2)   reference price 4.4890595342...

The goldens (price, per-step lstsq ranks, exercise-step histogram) and every path's exercise step
(tests/golden/big_exercise_steps.npz) were produced by the UNMODIFIED reference in the build container
(tests/golden/make_golden.py --with-big).  The injected normals are the reference's own seed-42 legacy stream,
regenerated here (np.random.normal is deterministic), so nothing minute-long runs on the CPU at test time.

Bars:
  * FP64 storage: 1e-10 relative and ZERO flipped decisions against the reference run.
  * FP32 path storage, double state: the SAME bar against the reference run on its own paths rounded to float32
    (`*_f32paths` goldens: the reference fed paths.astype(float32)) -- float storage changes the inputs, not the
    arithmetic.  Measured: identical to 15 digits, 0 flips.
  * The distance between the float-input and the double-input reference runs is the REFERENCE's sensitivity to a 6e-8
    relative input perturbation: exercise decisions are discontinuous in the paths, a fraction f ~ 7e-4 of the paths
    flips (each by an O(1) realised cashflow) and the price moves by ~2 sqrt(f / P) / price -- 1.45e-5 at 1M x 252,
    4.8e-7 at 500k x 100.  The 1e-5 FP32 tolerance of the north star is therefore a statement about the configuration's
    real size (100M paths: 1.4e-6 by the 1/sqrt(P) law, measured in tests/test_gpu_fp32_scaling.py), and is asserted
    here scaled by sqrt(P_full / P); flips and the budget used at full size are reported (-rP, gpurun_out/).
  * FP32 state on top of FP32 paths: same scaled bar against the float-input reference run.
"""
import json
import os

import numpy as np
import pytest

from oracle import lsm_oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REPORT = os.path.join(os.path.dirname(HERE), "gpurun_out", "big_shape_parity.jsonl")


@pytest.fixture(scope="module")
def big_steps():
    return np.load(os.path.join(HERE, "golden", "big_exercise_steps.npz"))


@pytest.fixture(scope="module")
def normals_cache():
    cache = {}

    def get(c):
        if c["name"] not in cache:
            cache.clear()                                      # one 2 GB array at a time
            np.random.seed(c["seed"])
            cache[c["name"]] = orc.draw_normals(c["n_paths"], c["n_time_steps"])
        return cache[c["name"]]
    return get


def _record(rec):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    print("BIG_SHAPE_PARITY " + json.dumps(rec))


MODES = [("float64", "float64"), ("float32", "float64"), ("float32", "float32")]
FULL_SIZE = {"c3_reduced": 100_000_000, "c5_reduced": 50_000_000}       # BASELINE.json configs[2], configs[4]


@pytest.mark.parametrize("name", ["c3_reduced", "c5_reduced"])
@pytest.mark.parametrize("path_dtype,state_dtype", MODES)
def test_reference_shapes_injected_normals(amc, golden, big_steps, normals_cache, name, path_dtype, state_dtype):
    c = golden[name]
    n, P = c["n_time_steps"], c["n_paths"]
    Z = normals_cache(c)
    dp = amc.paths_from_normals(Z, c["S0"], c["r"], c["sigma"], c["T"], dtype=path_dtype)
    dt = c["T"] / n
    res = amc.lsm_price(dp, c["K"], c["r"], dt, c["option_type"], c["barrier_level"], c["exercise_type"],
                        c["basis_type"], c["degree"], **c["kwargs"], want_exercise_steps=True, state_dtype=state_dtype)
    dp.free()
    fp64 = path_dtype == "float64"
    same_inputs = golden[name if fp64 else name + "_f32paths"]         # the reference run on the inputs the GPU holds
    tau_same = big_steps[same_inputs["name"]].astype(np.int32)
    flips_same = int((res.exercise_steps != tau_same).sum())
    err_same = abs(float(res.price) - same_inputs["price"]) / same_inputs["price"]
    flips_f64 = int((res.exercise_steps != big_steps[name].astype(np.int32)).sum())
    err_f64 = abs(float(res.price) - c["price"]) / c["price"]
    scale = (FULL_SIZE[name] / P) ** 0.5
    rec = dict(case=name, paths=P, steps=n, path_dtype=path_dtype, state_dtype=state_dtype, price=float(res.price),
               reference_same_inputs=same_inputs["price"], rel_err_same_inputs=err_same, flips_same_inputs=flips_same,
               reference_f64_inputs=c["price"], rel_err_vs_f64_inputs=err_f64, flips_vs_f64_inputs=flips_f64,
               flipped_fraction_vs_f64=flips_f64 / P, ranks_equal=res.rank[:n].tolist() == same_inputs["ranks"],
               fp32_budget_used_at_this_size=err_f64 / 1e-5, fp32_budget_used_at_full_size=err_f64 / scale / 1e-5,
               max_pivot_loss=float(np.max(res.pivot_loss)))
    _record(rec)
    if state_dtype == "float64":
        # same inputs, same arithmetic: the FP64 bar
        assert err_same <= 1e-10
        assert flips_same == 0, f"{flips_same} paths exercise at a different step than in the reference"
        assert np.bincount(res.exercise_steps, minlength=n + 1).tolist() == same_inputs["exercise_step_hist"]
        assert rec["ranks_equal"]
    if not fp64:
        assert err_f64 <= 1e-5 * scale                      # the FP32 tolerance by the 1/sqrt(P) law of flip noise
        assert rec["fp32_budget_used_at_full_size"] < 0.5
        assert flips_f64 / P < 5e-3
