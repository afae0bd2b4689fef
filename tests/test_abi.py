"""CPU tier: libamc.so builds for sm_100a, loads without a GPU, exports every symbol include/amc.h declares,
and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "amc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for must in ["amc_ctx_create", "amc_paths_generate", "amc_paths_from_normals", "amc_paths_from_host",
                 "amc_lsm_price", "amc_continuation", "amc_comm_init", "amc_intrinsic_value", "amc_regression_fit"]:
        assert must in syms


def test_library_exports_every_declared_symbol(libamc_path):
    lib = ctypes.CDLL(libamc_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol(libamc_path):
    from american_monte_carlo_b200 import _native
    assert sorted(_native.PROTOTYPES) == declared_symbols()
    _native.lib()


def test_library_is_sm100a_and_has_no_cpu_path(libamc_path):
    out = subprocess.run(["cuobjdump", "--list-elf", libamc_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure mode below only exists on a CPU box")
    import american_monte_carlo_b200 as pkg
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pkg.Context(0)
    import numpy as np
    pkg.set_default_context(None)
    with pytest.raises(RuntimeError):
        pkg.intrinsic_value(np.array([90.0, 100.0]), 100.0, "Put")


def test_unknown_basis_is_a_value_error_before_any_device_work():
    import american_monte_carlo_b200 as pkg
    import numpy as np
    with pytest.raises(ValueError, match="Unknown basis type 'Hermite'"):
        pkg.get_basis_polynomials(np.ones(3), "Hermite", 2)
    with pytest.raises(TypeError, match="unexpected keyword argument 'bogus'"):
        pkg.lsmc_option_pricing(np.ones((4, 3)), 1.0, 0.0, 0.1, "Put", bogus=1)


def test_shard_range_partitions_the_path_axis():
    from american_monte_carlo_b200 import shard_range
    for P, W in [(100_000_000, 8), (10, 3), (7, 8), (0, 2), (1, 1)]:
        parts = [shard_range(P, W, r) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == P
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_shard_range_in_quads_for_the_float_generator():
    """unit=4: every shard starts on a quad boundary (one Philox call of the float generator serves four adjacent paths;
    the path-free mode requires it), sizes differ by at most one quad, the union is the whole range."""
    from american_monte_carlo_b200 import shard_range
    for P, W in [(100_000_000, 8), (1_000_003, 2), (1_000_003, 3), (5, 4), (0, 2), (4, 8)]:
        parts = [shard_range(P, W, r, 4) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == P
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(lo % 4 == 0 for lo, hi in parts if hi > lo)
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 7                     # one quad, plus the ragged tail of the last shard


def test_path_free_mode_rejects_what_it_cannot_regenerate():
    import american_monte_carlo_b200 as pkg
    for kw in (dict(rng="numpy", dtype="float32"), dict(rng="philox", dtype="float64")):
        with pytest.raises(ValueError, match="store_paths=False"):
            pkg.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 10, 100, store_paths=False, **kw)


def test_c_example_builds_against_the_header_and_fails_loudly_without_a_gpu(libamc_path, tmp_path):
    """examples/price_put.c is a plain-C host of include/amc.h: it must compile and link against libamc.so; run without
    a CUDA device it must stop at amc_ctx_create with the library's message (no CPU path), with one it prices."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    exe = tmp_path / "price_put"
    libdir = os.path.dirname(libamc_path)
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "price_put.c"),
           "-L", libdir, "-l:libamc.so", "-lm", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    b = subprocess.run(cmd, capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    if r.returncode == 0:
        assert "American put K=40" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr


# the reference's positional parameter order (american_monte_carlo.py:72,85,90,98,110,126,139-140,171,180-182); a
# reference-style positional call must bind the same way through the shim
REFERENCE_SIGNATURES = {
    "generate_asset_paths": ["S0", "r", "sigma", "T", "n_time_steps", "n_paths"],
    "intrinsic_value": ["S", "K", "option_type"],
    "apply_exercise": ["cashflows", "exercise_times", "in_the_money_idx", "exercise_value", "continuation_estimated", "t"],
    "get_basis_polynomials": ["X", "basis_type", "degree"],
    "regression_estimate": ["X", "Y", "basis_type", "degree", "scaling", "scaling_factor"],
    "estimate_continuation_values": ["paths", "t", "r", "dt", "cashflows", "exercise_times", "basis_type", "degree"],
    "perform_backward_iteration": ["K", "r", "dt", "n_time_steps", "barrier_hit", "cashflows", "paths", "option_type",
                                   "exercise_times", "exercise_type", "continuation_values", "basis_type", "degree"],
    "precompute_barrier_hit_matrix": ["paths", "barrier_level"],
    "lsmc_option_pricing": ["paths", "K", "r", "dt", "option_type", "barrier_level", "exercise_type", "basis_type", "degree"],
    "compute_ccr_exposures": ["continuation_values"],
}


def _positional(fn):
    import inspect
    return [p.name for p in inspect.signature(fn).parameters.values()
            if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]


def test_shim_keeps_the_reference_positional_order():
    import american_monte_carlo as shim
    for name, want in REFERENCE_SIGNATURES.items():
        assert _positional(getattr(shim, name)) == want, name


def test_signature_table_matches_the_reference_when_it_is_present():
    ref_py = "/root/reference/american_monte_carlo.py"
    if not os.path.exists(ref_py):
        pytest.skip("reference checkout not present (GPU box)")
    import ast
    tree = ast.parse(open(ref_py).read())
    defs = {n.name: [a.arg for a in n.args.args] for n in tree.body if isinstance(n, ast.FunctionDef)}
    for name, want in REFERENCE_SIGNATURES.items():
        assert defs[name] == want, name
