"""CPU tier: the bench.py contract that can be checked without a GPU -- the reference arm (the oracle timed on the host)
prints exactly one JSON line with the agreed keys -- and the pure host logic of the sweep drivers."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "path-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "c1"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "paths x 50 steps" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 1e5 < d["value"] < 1e9                       # a NumPy port on a handful of cores
    assert abs(cb["price"] - 4.4783398987704475) < 1e-12   # config 1, seed 42: the reference's own number


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_contract_grid_definition_and_cell_sharding():
    sys.path.insert(0, ROOT)
    from american_monte_carlo_b200 import sweeps
    strikes, vols, mats = sweeps.default_contract_grid()
    assert len(strikes) * len(vols) * len(mats) == 1024
    assert strikes[0] == 32.0 and strikes[-1] == 48.0 and abs(vols[0] - 0.10) < 1e-15 and mats[-1] == 2.0
    cells = len(vols) * len(mats)
    for world in (1, 2, 4, 8, 3):
        owned = [sweeps.cells_of_rank(cells, world, r) for r in range(world)]
        assert sorted(c for o in owned for c in o) == list(range(cells))           # a partition
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1        # balanced
