import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import american_monte_carlo_b200 as amc
from oracle import lsm_oracle as orc
rng = np.random.default_rng(2024)
bad = []
losses_ok, losses_bad = [], []
import warnings
warnings.simplefilter("ignore")
N = 300
for it in range(N):
    P = int(rng.integers(200, 4000)); n = int(rng.integers(2, 12))
    S0 = float(rng.choice([20.0, 36.0, 100.0, 250.0])); K = S0 * float(rng.uniform(0.85, 1.15))
    r = float(rng.uniform(0, 0.08)); sigma = float(rng.uniform(0.05, 0.6)); T = float(rng.choice([0.25, 1.0, 3.0]))
    basis = str(rng.choice(["Power", "Chebyshev", "Legendre"])); degree = int(rng.integers(4, 11))
    scaling = bool(rng.integers(0, 2)); opt = str(rng.choice(["Put", "Call"]))
    kw = dict(scaling=True, scaling_factor=float(rng.choice([1, 2]))) if scaling else {}
    np.random.seed(it)
    Z = orc.draw_normals(P, n)
    paths = orc.paths_from_normals(Z, S0, r, sigma, T)
    want = orc.lsm_backward(paths, K, r, T / n, opt, None, "American", basis, degree, keep_continuation=False, keep_diag=True, **kw)
    dp = amc.paths_from_normals(Z, S0, r, sigma, T)
    got = amc.lsm_price(dp, K, r, T / n, opt, None, "American", basis, degree, want_exercise_steps=True, **kw)
    dp.free()
    flips = int((got.exercise_steps != want.exercise_times).sum())
    ranks_o = [want.steps[t]["rank"] for t in range(n)]
    ranks_g = got.rank[:n].tolist()
    rel = abs(got.price - want.price) / max(abs(want.price), 1e-9)
    loss = float(got.pivot_loss.max())
    (losses_bad if (flips or ranks_o != ranks_g or rel > 1e-9) else losses_ok).append(loss)
    if flips or ranks_o != ranks_g or rel > 1e-9:
        # how close to the cutoff were the steps where ranks differ?
        near = []
        for t in range(n):
            if ranks_o[t] != ranks_g[t]:
                sv = want.steps[t]["sv"]; cut = np.finfo(float).eps * max(P, degree + 1) * sv[0]
                near.append((t, ranks_o[t], ranks_g[t], float(min(abs(np.array(sv) / cut - 1)))))
        bad.append(dict(loss=loss, it=it, P=P, n=n, basis=basis, degree=degree, scaling=scaling, flips=flips, rel=rel, near=near))
print(json.dumps(dict(total=N, bad=len(bad), cases=bad[:20]), indent=1))
lo = np.sort(np.array(losses_ok))
print("ok cases: max loss %.3g, 99th pct %.3g, median %.3g" % (lo[-1], lo[int(0.99 * len(lo))], lo[len(lo) // 2]))
print("top ok losses", lo[-8:])
print("bad losses", losses_bad)
