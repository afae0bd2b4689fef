"""Generate tests/golden/golden.json from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden.py [--with-c2]

* imports /root/reference/american_monte_carlo.py with matplotlib / QuantLib stubbed (they are
  not installed here and are not on the hot path);
* for every case runs the reference's own generate_asset_paths + lsmc_option_pricing, wrapping
  np.linalg.lstsq to capture numpy's rank / singular values / coefficients per backward step;
* runs oracle/lsm_oracle.py on the same seed and REFUSES to write the file unless the oracle is
  bit-identical to the reference (paths, price, every continuation vector);
* stores prices at full precision plus fingerprints of the exercise decisions (histogram of
  exercise steps, sum of cashflows) so a GPU parity test can report flipped decisions.

`--with-c2` adds BASELINE.json configs[1] (10M paths x 50 steps, ~4 min, ~30 GB RSS).  Without
the flag an existing c2 entry in golden.json is carried over unchanged.
`--with-big` adds the reduced shapes of configs[2] and configs[4] that SURVEY.md section 8(c) pins
(`c3_reduced` 1M x 252 Power-3 -> 4.4847469992..., `c5_reduced` 500k x 100 Legendre-8 scaled ->
4.4890595342...; ~10 min, ~25 GB RSS) and writes every path's exercise step of those two runs to
tests/golden/big_exercise_steps.npz (uint8, compressed) so that the GPU tier can count flipped
decisions exactly without re-running a minute-long CPU sweep.  Carried over like c2 otherwise.
/root/reference does not exist on the GPU box; nothing under tests/ reads it at test time.
"""
import argparse
import json
import os
import sys
import time
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

for _m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
           "matplotlib.gridspec", "QuantLib"]:
    sys.modules[_m] = MagicMock()
sys.path.insert(0, "/root/reference")
import american_monte_carlo as ref  # noqa: E402  (the real reference)

from oracle import lsm_oracle as orc  # noqa: E402

LS = dict(S0=36.0, K=40.0, r=0.06, sigma=0.2, T=1.0)          # Longstaff-Schwartz Table 1 put
NB = dict(S0=95, K=100, r=0.01, sigma=0.2, T=1.0)             # notebook cell 5
UT = dict(S0=100, K=100, r=0.01, sigma=0.2, T=1.0)            # unit_test.py:45


def case(name, mkt, n, P, option_type, exercise_type, barrier, basis, degree, kwargs=None,
         seed=42, steps_detail=False, source="", round_paths_f32=False):
    return dict(name=name, **mkt, n_time_steps=n, n_paths=P, option_type=option_type,
                exercise_type=exercise_type, barrier_level=barrier, basis_type=basis, degree=degree,
                kwargs=kwargs or {}, seed=seed, steps_detail=steps_detail, source=source,
                round_paths_f32=round_paths_f32)


CASES = []
nbkw = dict(scaling=True, scaling_factor=1)
CASES += [
    case("nb_european_put", NB, 100, 1000, "Put", "European", None, "Chebyshev", 10, nbkw,
         source="AmericanMonteCarlo.ipynb:202-203 (prints 9.9485)"),
    case("nb_american_put", NB, 100, 1000, "Put", "American", None, "Chebyshev", 10, nbkw,
         steps_detail=True, source="AmericanMonteCarlo.ipynb:248-249 (prints 10.3838)"),
    case("nb_di70_european_put", NB, 100, 1000, "Put", "European", 70, "Chebyshev", 10, nbkw,
         source="AmericanMonteCarlo.ipynb:312-313 (prints 4.0108)"),
    case("nb_di70_european_put_200x10000", NB, 200, 10000, "Put", "European", 70, "Chebyshev", 10, nbkw,
         source="AmericanMonteCarlo.ipynb:377-378 (prints 3.7835)"),
    case("nb_di70_european_put_unscaled", NB, 100, 1000, "Put", "European", 70, "Chebyshev", 10,
         dict(scaling=False, scaling_factor=1), source="AmericanMonteCarlo.ipynb:442-443 (prints 4.0108)"),
]
NB_PRINTED = {"nb_european_put": "9.9485", "nb_american_put": "10.3838", "nb_di70_european_put": "4.0108",
              "nb_di70_european_put_200x10000": "3.7835", "nb_di70_european_put_unscaled": "4.0108"}
for ot, et, pct in [("Put", "European", None), ("Call", "European", None), ("Put", "American", None),
                    ("Call", "American", None), ("Put", "European", 80), ("Call", "European", 80),
                    ("Put", "American", 80), ("Call", "American", 80), ("Put", "European", 60),
                    ("Call", "European", 60), ("Put", "American", 60), ("Call", "American", 60)]:
    b = UT["S0"] * pct / 100 if pct else None                  # unit_test.py:47
    CASES.append(case(f"ut_{ot}_{et}_{pct}", UT, 100, 10000, ot, et, b, "Chebyshev", 4,
                      steps_detail=(et == "American"), source="unit_test.py:30-50"))
CASES += [
    case("c1_power3", LS, 50, 100000, "Put", "American", None, "Power", 3, steps_detail=True,
         source="BASELINE.json configs[0]"),
    case("c1_chebyshev3", LS, 50, 100000, "Put", "American", None, "Chebyshev", 3),
    case("c1_legendre3", LS, 50, 100000, "Put", "American", None, "Legendre", 3),
    case("c1_call_power3", LS, 50, 100000, "Call", "American", None, "Power", 3),
    case("c1_di30_power3", LS, 50, 100000, "Put", "American", 30.0, "Power", 3),
    case("small_power8_unscaled", LS, 25, 20000, "Put", "American", None, "Power", 8, steps_detail=True,
         source="rank-truncated degree-8 stress (SURVEY.md 0.3)"),
    case("small_legendre8_scaled", LS, 25, 20000, "Put", "American", None, "Legendre", 8,
         dict(scaling=True, scaling_factor=2), steps_detail=True, source="config 5 shape, reduced"),
    case("small_chebyshev10_unscaled", NB, 20, 5000, "Put", "American", None, "Chebyshev", 10,
         steps_detail=True),
    case("small_degree0", LS, 10, 5000, "Put", "American", None, "Power", 0, steps_detail=True),
    case("small_degree1_scaled", LS, 10, 5000, "Put", "American", None, "Chebyshev", 1,
         dict(scaling=True), steps_detail=True),
    case("deep_itm_exercise_at_0", dict(S0=30.0, K=40.0, r=0.06, sigma=0.2, T=1.0), 20, 5000, "Put",
         "American", None, "Power", 3, source="SURVEY.md 0.2: price == intrinsic(S0) == 10"),
    case("tiny_paths_lt_k", LS, 5, 3, "Put", "American", None, "Power", 3, steps_detail=True),
]
C2 = case("c2_power3_10M", LS, 50, 10_000_000, "Put", "American", None, "Power", 3, steps_detail=True,
          source="BASELINE.json configs[1]")


BIG = [
    case("c3_reduced", LS, 252, 1_000_000, "Put", "American", None, "Power", 3, steps_detail=True,
         source="BASELINE.json configs[2] at 1M paths (SURVEY.md 8c: 4.4847469992)"),
    case("c5_reduced", LS, 100, 500_000, "Put", "American", None, "Legendre", 8,
         dict(scaling=True, scaling_factor=2), steps_detail=True,
         source="BASELINE.json configs[4] at 500k paths, the reference-comparable basis (SURVEY.md 8c: 4.4890595342)"),
]
# The same two runs with the reference fed its own paths ROUNDED TO FLOAT32 (and widened back): what "FP32 path
# storage" means as an input to the unmodified reference.  The CUDA path with float storage must reproduce THESE to the
# FP64 bar (same arithmetic, same inputs); their distance from the unrounded runs is the reference's own sensitivity to
# a 6e-8 relative perturbation of its inputs (exercise decisions are discontinuous in the paths).
BIG += [
    case("c3_reduced_f32paths", LS, 252, 1_000_000, "Put", "American", None, "Power", 3, steps_detail=True,
         source="c3_reduced with paths.astype(float32).astype(float64)", round_paths_f32=True),
    case("c5_reduced_f32paths", LS, 100, 500_000, "Put", "American", None, "Legendre", 8,
         dict(scaling=True, scaling_factor=2), steps_detail=True,
         source="c5_reduced with paths.astype(float32).astype(float64)", round_paths_f32=True),
]
BIG_EXPECT = {"c3_reduced": "4.4847469992", "c5_reduced": "4.4890595342"}


class LstsqTap:
    """Wrap np.linalg.lstsq to record what numpy itself reports for each call."""

    def __init__(self):
        self.calls = []
        self._orig = np.linalg.lstsq

    def __enter__(self):
        def tapped(A, Y, rcond=None):
            out = self._orig(A, Y, rcond=rcond)
            self.calls.append(dict(rank=int(out[2]), sv=[float(s) for s in out[3]],
                                   coeffs=[float(c) for c in out[0]]))
            return out
        np.linalg.lstsq = tapped
        return self

    def __exit__(self, *a):
        np.linalg.lstsq = self._orig


def run_case(c, check_oracle=True, keep_tau=None):
    t0 = time.time()
    mk = (c["S0"], c["r"], c["sigma"], c["T"], c["n_time_steps"], c["n_paths"])
    dt = c["T"] / c["n_time_steps"]
    np.random.seed(c["seed"])
    paths = ref.generate_asset_paths(*mk)
    if c.get("round_paths_f32"):
        paths = paths.astype(np.float32).astype(np.float64)
    with LstsqTap() as tap:
        price, cont = ref.lsmc_option_pricing(paths, c["K"], c["r"], dt, c["option_type"], c["barrier_level"],
                                              c["exercise_type"], c["basis_type"], c["degree"], **c["kwargs"])
    # the reference does not return its exercise state: recover it by calling its own pieces
    n, P = c["n_time_steps"], c["n_paths"]
    cf = np.zeros(P)
    tau = np.full(P, n)
    ref.perform_backward_iteration(c["K"], c["r"], dt, n, ref.precompute_barrier_hit_matrix(paths, c["barrier_level"]),
                                   cf, paths, c["option_type"], tau, c["exercise_type"], [], c["basis_type"],
                                   c["degree"], **c["kwargs"])
    assert np.mean(cf * np.exp(-c["r"] * dt * tau)) == price

    if check_oracle:
        np.random.seed(c["seed"])
        o_paths = orc.generate_asset_paths(*mk)
        if c.get("round_paths_f32"):
            o_paths = o_paths.astype(np.float32).astype(np.float64)
        assert np.array_equal(o_paths, paths), c["name"]
        o = orc.lsm_backward(o_paths, c["K"], c["r"], dt, c["option_type"], c["barrier_level"], c["exercise_type"],
                             c["basis_type"], c["degree"], keep_diag=True, **c["kwargs"])
        assert o.price == price, (c["name"], o.price, price)
        assert np.array_equal(o.cashflows, cf) and np.array_equal(o.exercise_times, tau), c["name"]
        assert len(o.continuation_values) == len(cont) == n + 1
        for (ta, sa, ca), (tb, sb, cb) in zip(o.continuation_values, cont):
            assert ta == tb and np.array_equal(sa, sb) and np.array_equal(ca, cb), (c["name"], ta)
        # tap order is t = n-1 .. 0
        for i, call in enumerate(tap.calls):
            t = n - 1 - i
            assert o.steps[t]["rank"] == call["rank"]
        del o, o_paths

    rec = {k: c[k] for k in c if k != "steps_detail"}
    rec["price"] = float(price)
    rec["cashflow_sum"] = float(np.sum(cf))
    rec["exercise_step_hist"] = np.bincount(tau, minlength=n + 1).tolist()
    rec["n_nonzero_cashflows"] = int(np.count_nonzero(cf))
    rec["ranks"] = [tap.calls[n - 1 - t]["rank"] for t in range(n)]          # indexed by t = 0..n-1
    rec["path_checksum"] = [float(paths[:, n].sum()), float(paths[:, n // 2].sum()), float(paths.min()),
                            float(paths.max())]
    if c["steps_detail"]:
        pick = sorted(set([0, 1, 2, n // 2, n - 2, n - 1]) & set(range(n)))
        rec["steps"] = {str(t): tap.calls[n - 1 - t] for t in pick}
        # fitted continuation value at a few probe paths (first 8) for those steps
        rec["cont_probe"] = {str(t): [float(v) for v in cont[t][2][:8]] for t in pick}
        rec["cont_mean"] = {str(t): float(np.mean(cont[t][2])) for t in pick}
    rec["seconds"] = round(time.time() - t0, 2)
    if keep_tau is not None:
        assert n < 256
        keep_tau[c["name"]] = tau.astype(np.uint8)
    if c["name"] in BIG_EXPECT:
        assert f"{price:.10f}" == BIG_EXPECT[c["name"]], (c["name"], price)
    if c["name"] in NB_PRINTED:
        assert f"{price:.4f}" == NB_PRINTED[c["name"]], (c["name"], price)
        rec["printed_in_notebook"] = NB_PRINTED[c["name"]]
    print(f"{c['name']:36s} price={price!r:22} ranks(min/max)={min(rec['ranks'] or [0])}/{max(rec['ranks'] or [0])}"
          f"  {rec['seconds']}s", flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--with-c2", action="store_true")
    ap.add_argument("--with-big", action="store_true")
    ap.add_argument("--only-big", action="store_true", help="carry every other case over from the existing file")
    ap.add_argument("--big-names", default="", help="comma-separated subset of the big cases to (re)run; the others carry over")
    args = ap.parse_args()
    out_path = os.path.join(HERE, "golden.json")
    old = {}
    if os.path.exists(out_path):
        old = {r["name"]: r for r in json.load(open(out_path))["cases"]}
    if args.only_big:
        recs = [old[c["name"]] for c in CASES]
        args.with_big = True
    else:
        recs = [run_case(c) for c in CASES]
    if args.with_c2:
        recs.append(run_case(C2))
    elif C2["name"] in old:
        recs.append(old[C2["name"]])
    if args.with_big:
        npz = os.path.join(HERE, "big_exercise_steps.npz")
        taus = dict(np.load(npz)) if os.path.exists(npz) else {}
        only = set(filter(None, args.big_names.split(",")))
        for c in BIG:
            if only and c["name"] not in only and c["name"] in old and c["name"] in taus:
                recs.append(old[c["name"]])
            else:
                recs.append(run_case(c, keep_tau=taus))
        np.savez_compressed(npz, **taus)
    else:
        recs += [old[c["name"]] for c in BIG if c["name"] in old]
    meta = dict(numpy=np.__version__, generator="tests/golden/make_golden.py",
                reference="/root/reference/american_monte_carlo.py (unmodified; matplotlib/QuantLib stubbed)",
                note="oracle/lsm_oracle.py was asserted bit-identical to the reference on every case")
    with open(out_path, "w") as f:
        json.dump(dict(meta=meta, cases=recs), f, indent=1)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
