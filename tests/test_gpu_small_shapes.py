"""GPU tier: every kernel on small, ragged shapes (path counts that are not multiples of the tile, vector or warp
size; 3..12 steps; both storage types; barrier / exercise-step / SVD / batch / exposures variants).  compute-sanitizer
is not available on the GPU pool, so this walk -- with every result checked finite and shapes checked -- plus the oracle
parity tests is what guards the indexing."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_kernels_on_ragged_small_shapes(amc):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sanitize_small
    sanitize_small.main()


def test_persistent_sweep_on_stored_sets_stays_correct(libamc_path):
    """AMC_PERSISTENT=1 (opt-in: the cooperative one-launch sweep kernel on STORED path sets; the default for them is the
    launch chain) must keep giving the oracle's price and decisions.  The switch is read once per process, hence the
    subprocess."""
    import subprocess
    code = (
        "import numpy as np, american_monte_carlo_b200 as amc\n"
        "from oracle import lsm_oracle as orc\n"
        "np.random.seed(3); Z = orc.draw_normals(30011, 20)\n"
        "paths = orc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0)\n"
        "for dtype, kw in (('float64', {}), ('float64', dict(barrier=33.0)), ('float64', dict(ex='European'))):\n"
        "    want = orc.lsm_backward(paths, 40.0, 0.06, 0.05, 'Put', kw.get('barrier'), kw.get('ex', 'American'), 'Power', 3, keep_continuation=False)\n"
        "    dp = amc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0, dtype=dtype)\n"
        "    res = amc.lsm_price(dp, 40.0, 0.06, 0.05, 'Put', kw.get('barrier'), kw.get('ex', 'American'), 'Power', 3, want_exercise_steps=True)\n"
        "    assert res.timing['step_launches'] == 1, res.timing\n"
        "    assert int((res.exercise_steps != want.exercise_times).sum()) == 0\n"
        "    assert abs(res.price - want.price) <= 1e-10 * want.price\n"
        "print('persistent ok')\n")
    env = dict(os.environ, AMC_PERSISTENT="1", PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0 and "persistent ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


# ---- one-cluster sweep kernel (lsm_cluster.cuh): what prices small stored sets by default ------------------------------
CLUSTER_CASES = [
    # name, P, n, dtype, state, contract kwargs, basis, degree, lsm kwargs
    ("put_f64", 30011, 20, "float64", "float64", dict(), "Power", 3, dict()),
    ("put_barrier_f64", 30011, 20, "float64", "float64", dict(barrier=33.0), "Power", 3, dict()),
    ("european_f64", 30011, 20, "float64", "float64", dict(ex="European"), "Chebyshev", 4, dict()),
    ("call_scaled_f64", 12345, 7, "float64", "float64", dict(opt="Call", K=34.0), "Legendre", 6, dict(scaling=True)),
    ("put_deg8_scaled_f64", 20000, 10, "float64", "float64", dict(), "Laguerre", 8, dict(scaling=True)),
    ("put_deg0_f64", 999, 5, "float64", "float64", dict(), "Power", 0, dict()),
    ("put_deg1_f64", 31, 3, "float64", "float64", dict(), "Power", 1, dict()),
    ("put_f32paths", 30011, 20, "float32", "float64", dict(), "Power", 3, dict()),
    ("put_f32state", 30011, 20, "float32", "float32", dict(), "Power", 3, dict()),
    ("put_cheb4_unscaled_f64", 25000, 12, "float64", "float64", dict(), "Chebyshev", 4, dict()),   # rank-truncated steps
    # the remaining degrees the cluster kernel is instantiated for (7 and 16 accumulators: other reduction widths)
    ("put_deg2_f64", 4097, 9, "float64", "float64", dict(), "Power", 2, dict()),
    ("put_deg5_scaled_f64", 33333, 8, "float64", "float64", dict(), "Legendre", 5, dict(scaling=True)),
    ("call_deg5_barrier_f32state", 8191, 6, "float32", "float32", dict(opt="Call", K=35.0, barrier=38.0), "Chebyshev", 5,
     dict(scaling=True)),
    ("european_deg2_f32paths", 2049, 5, "float32", "float64", dict(ex="European"), "Legendre", 2, dict()),
    # one and two time steps: no column is fetched ahead / exactly one is
    ("put_one_step", 5000, 1, "float64", "float64", dict(), "Power", 3, dict()),
    ("put_two_steps", 5000, 2, "float64", "float64", dict(), "Power", 3, dict()),
]

_CLUSTER_WORKER = r"""
import sys, numpy as np, american_monte_carlo_b200 as amc
sys.path.insert(0, %(tests)r)
from test_gpu_small_shapes import CLUSTER_CASES
out = {}
for name, P, n, dtype, state, ckw, basis, deg, kw in CLUSTER_CASES:
    Z = np.random.default_rng(P + n).standard_normal((P, n))
    dp = amc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0, dtype=dtype)
    r = amc.lsm_price(dp, ckw.get('K', 40.0), 0.06, 1.0 / n, ckw.get('opt', 'Put'), ckw.get('barrier'), ckw.get('ex', 'American'),
                      basis, deg, want_exercise_steps=True, want_cashflows=True, want_regression=True, state_dtype=state, **kw)
    out[name + '.price'] = np.float64(r.price)
    out[name + '.kind'] = np.int64(r.timing['sweep_kind'])
    out[name + '.steps'] = r.exercise_steps
    out[name + '.cash'] = r.cashflow0
    out[name + '.gamma'] = r.gamma
    out[name + '.rank'] = r.rank
    dp.free()
np.savez(sys.argv[1], **out)
print('worker ok')
"""


def _run_cluster_worker(tmp_path, tag, env_extra):
    import subprocess
    import numpy as np
    out = os.path.join(str(tmp_path), tag + ".npz")
    env = dict(os.environ, PYTHONPATH=ROOT, **env_extra)
    code = _CLUSTER_WORKER % dict(tests=os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-c", code, out], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0 and "worker ok" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
    return dict(np.load(out))


def test_cluster_kernel_equals_launch_chain_and_oracle(amc, tmp_path):
    """Small stored sets are priced by the one-cluster kernel (sweep_kind 2); AMC_CLUSTER=0 keeps them on the launch chain
    (sweep_kind 0).  Same decisions on every path, prices to rounding (the partial sums are grouped differently), same
    regression ranks, coefficient tables to 1e-9; and the cluster result against the oracle on the f64 cases."""
    import numpy as np
    from oracle import lsm_oracle as orc
    a = _run_cluster_worker(tmp_path, "cluster", {"AMC_CLUSTER": "1"})
    b = _run_cluster_worker(tmp_path, "chain", {"AMC_CLUSTER": "0"})
    for name, P, n, dtype, state, ckw, basis, deg, kw in CLUSTER_CASES:
        assert int(a[name + ".kind"]) == (2 if deg <= 5 else 0), (name, "degrees 0..5 are priced by the cluster kernel")
        assert int(b[name + ".kind"]) == 0, name
        flips = int((a[name + ".steps"] != b[name + ".steps"]).sum())
        assert flips == 0, (name, flips)
        pa, pb = float(a[name + ".price"]), float(b[name + ".price"])
        tol = 1e-12 if state == "float64" else 1e-7
        assert abs(pa - pb) <= tol * abs(pb), (name, pa, pb)
        np.testing.assert_array_equal(a[name + ".rank"], b[name + ".rank"], err_msg=name)
        if state == "float64":
            np.testing.assert_array_equal(a[name + ".cash"], b[name + ".cash"], err_msg=name)
        if bool((b[name + ".rank"][:n] == deg + 1).all()):       # rank-truncated fits amplify rounding: decisions above
            np.testing.assert_allclose(a[name + ".gamma"], b[name + ".gamma"], rtol=1e-7, atol=1e-9, err_msg=name)
        if dtype == "float64":
            Z = np.random.default_rng(P + n).standard_normal((P, n))
            paths = orc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0)
            want = orc.lsm_backward(paths, ckw.get("K", 40.0), 0.06, 1.0 / n, ckw.get("opt", "Put"), ckw.get("barrier"),
                                    ckw.get("ex", "American"), basis, deg, keep_continuation=False, **kw)
            assert int((a[name + ".steps"] != want.exercise_times).sum()) == 0, name
            assert abs(pa - want.price) <= 1e-10 * abs(want.price), (name, pa, want.price)


def test_cluster_kernel_capacity_edge(libamc_path):
    """Path counts around what one cluster's shared memory holds (AMC_CLUSTER_MAX_PATHS lifts the default cut at 147456
    paths, which is a speed threshold, not a capacity): the largest set the cluster kernel takes and the first one that
    goes to the launch chain agree with the oracle alike (3 steps keep the oracle quick).  Where the capacity ends depends
    on the cluster the device grants (16 CTAs: 145408 f64 paths at degree 3; 8 CTAs: half of that) -- the test asks for one
    threshold somewhere between the first and the last size."""
    import subprocess
    code = (
        "import numpy as np, american_monte_carlo_b200 as amc\n"
        "from oracle import lsm_oracle as orc\n"
        "n = 3; kinds = []\n"
        "for P in (60_000, 72_000, 74_000, 131_072, 145_408, 145_409, 150_000, 160_000):\n"
        "    Z = np.random.default_rng(P).standard_normal((P, n))\n"
        "    paths = orc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0)\n"
        "    want = orc.lsm_backward(paths, 40.0, 0.06, 1.0 / n, 'Put', None, 'American', 'Power', 3, keep_continuation=False)\n"
        "    dp = amc.paths_from_normals(Z, 36.0, 0.06, 0.2, 1.0)\n"
        "    res = amc.lsm_price(dp, 40.0, 0.06, 1.0 / n, 'Put', None, 'American', 'Power', 3, want_exercise_steps=True)\n"
        "    kinds.append(res.timing['sweep_kind'])\n"
        "    assert int((res.exercise_steps != want.exercise_times).sum()) == 0, P\n"
        "    assert abs(res.price - want.price) <= 1e-10 * want.price, (P, res.price, want.price)\n"
        "    dp.free()\n"
        "assert kinds[0] == 2 and kinds[-1] == 0, kinds\n"
        "assert kinds == sorted(kinds, reverse=True), kinds\n"
        "print('edge ok', kinds)\n")
    env = dict(os.environ, AMC_CLUSTER_MAX_PATHS="100000000", PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0 and "edge ok" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]


def test_cluster_kernel_default_threshold(amc):
    """By default the cluster kernel takes sets of up to 147456 paths (where it is at least as fast as the chain on B200:
    profiles/r2_cluster_vs_chain.md); float paths, whose capacity in a 16-CTA cluster is beyond that, show the threshold."""
    kinds = []
    for P in (60_000, 147_000, 147_456, 147_457):
        dp = amc.generate_asset_paths(36.0, 0.06, 0.2, 1.0, 4, P, rng="philox", seed=11, dtype="float32")
        kinds.append(amc.lsm_price(dp, 40.0, 0.06, 0.25, "Put", None, "American", "Power", 3).timing["sweep_kind"])
        dp.free()
    assert kinds[0] == 2 and kinds[3] == 0 and kinds == sorted(kinds, reverse=True), kinds
    if kinds[1] == 2:                    # the capacity reaches the threshold (16-CTA cluster): the cut is exactly there
        assert kinds[2] == 2, kinds


def test_launch_chain_on_ragged_small_shapes(libamc_path):
    """The walk of test_all_kernels_on_ragged_small_shapes again with the cluster kernel switched off: the launch chain
    (and its CUDA-graph replay) stays covered on small shapes."""
    import subprocess
    env = dict(os.environ, AMC_CLUSTER="0", PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sanitize_small.py")], capture_output=True, text=True,
                       timeout=900, env=env, cwd=ROOT)
    assert p.returncode == 0 and "sanitize_small: ok" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
