"""GPU tier: how the FP32-storage price error scales with the number of paths.

Float storage perturbs every path value by <= 6e-8 relative.  Exercise decisions are discontinuous in the paths, so a
fraction f of the paths flips its exercise step and each flip moves that path's realised cashflow by O(1): the price
difference between the float-storage and the double-storage sweep ON THE SAME NORMALS behaves like 2 sqrt(f / P) -- it
falls as 1/sqrt(P).  This test measures it at 252 steps for P = 0.25M .. 16M (normals generated on the device, the same
array feeding both storages) and checks the law, which puts the 100M-path configuration (BASELINE.json configs[2]) at
~1.5e-6, inside its 1e-5 tolerance, while 1M paths sit at ~1.5e-5 (tests/test_gpu_big_shapes.py).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_fp32_storage_error_falls_like_inverse_sqrt_paths(amc):
    import torch
    from american_monte_carlo_b200 import _native as N
    ctx = amc.default_context()
    S0, K, r, sigma, T, n = 36.0, 40.0, 0.06, 0.2, 1.0, 252
    rows = []
    for P in (250_000, 1_000_000, 4_000_000, 16_000_000):
        g = torch.Generator(device="cuda")
        g.manual_seed(1234)
        Z = torch.randn((P, n), dtype=torch.float64, device="cuda", generator=g)
        torch.cuda.synchronize()
        out = {}
        for name, did, state in (("f64", N.F64, "float64"), ("f32", N.F32, "float64"), ("f32s", N.F32, "float32")):
            h = C.c_void_p()
            N.check(N.lib().amc_paths_from_normals_dev(ctx.handle, Z.data_ptr(), S0, r, sigma, T, n, P, P, did, C.byref(h)))
            dp = amc.DevicePaths(ctx, h, P, P, n, did)
            res = amc.lsm_price(dp, K, r, T / n, "Put", None, "American", "Power", 3, want_exercise_steps=True,
                                state_dtype=state)
            out[name] = (float(res.price), res.exercise_steps.copy())
            dp.free()
        del Z
        torch.cuda.empty_cache()
        p64, t64 = out["f64"]
        row = dict(paths=P, steps=n, price_f64=p64)
        for name in ("f32", "f32s"):
            pr, tt = out[name]
            row[name + "_rel_err"] = abs(pr - p64) / p64
            row[name + "_flipped_fraction"] = float((tt != t64).mean())
            # what the flip-noise law predicts: 2 sqrt(f / P) cashflow units, relative to the price
            row[name + "_law"] = 2.0 * (row[name + "_flipped_fraction"] / P) ** 0.5 / p64
        rows.append(row)
        print("FP32_SCALING " + json.dumps(row))
    try:
        os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "fp32_scaling.json"), "w") as f:
            json.dump(rows, f, indent=1)
    except OSError:
        pass
    for row in rows:
        for name in ("f32", "f32s"):
            assert row[name + "_flipped_fraction"] < 3e-3
            assert row[name + "_rel_err"] < 4.0 * row[name + "_law"] + 1e-7        # within the law's noise band
    # the 16M-path point is already inside the FP32 tolerance; 100M is sqrt(6.25) = 2.5x closer still
    assert rows[-1]["f32_rel_err"] < 1e-5 and rows[-1]["f32s_rel_err"] < 1e-5
