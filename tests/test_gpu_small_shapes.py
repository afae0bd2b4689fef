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
