"""Test configuration.

Tiers:  `-m "not gpu"`  oracle vs golden vectors, host logic, solver source compiled for the host, C-ABI symbols
        `-m gpu`        parity tests proper: CUDA path (through the C ABI) vs oracle / golden, on a B200
Only tests (never the product) import `oracle/`.
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full-size configuration, minutes")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        data = json.load(f)
    return {c["name"]: c for c in data["cases"]}


@pytest.fixture(scope="session")
def libamc_path():
    """Build libamc.so if it is missing or older than its sources (nvcc cross-compiles without a GPU)."""
    from american_monte_carlo_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def amc(libamc_path):
    import american_monte_carlo_b200 as pkg
    return pkg


def oracle_case(case):
    """Run the CPU oracle on a golden case; returns (Z, paths, LsmResult)."""
    import numpy as np
    from oracle import lsm_oracle as orc
    np.random.seed(case["seed"])
    Z = orc.draw_normals(case["n_paths"], case["n_time_steps"])
    paths = orc.paths_from_normals(Z, case["S0"], case["r"], case["sigma"], case["T"])
    dt = case["T"] / case["n_time_steps"]
    res = orc.lsm_backward(paths, case["K"], case["r"], dt, case["option_type"], case["barrier_level"],
                           case["exercise_type"], case["basis_type"], case["degree"], keep_diag=True, **case["kwargs"])
    return Z, paths, res
