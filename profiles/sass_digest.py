"""Instruction mix of every kernel in libamc.so (cuobjdump -sass), the evidence for what each kernel runs on:
UBLKCP (1-D TMA bulk copies) / SYNCS (mbarrier) in the sweep kernels, UCGABAR (hardware cluster barrier) in the one-cluster kernel, LDGSTS (cp.async) in the injected-normals path
kernel, DFMA / DADD / DMUL (FP64 pipe), MUFU + I2FP/F2F (XU pipe) in the generators; no tensor-core mnemonics by design.

    python profiles/sass_digest.py > profiles/r2_sass_digest.txt          (build container, no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "american_monte_carlo_b200", "libamc.so")
WATCH = ["UBLKCP", "UTMALDG", "UCGABAR", "SYNCS", "LDGSTS", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "I2FP", "F2F", "FFMA", "IMAD.WIDE",
         "LOP3", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOM", "RED", "SHFL", "BAR", "UTCHMMA", "HMMA", "LDTM"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "UCGABAR" and op.startswith("UCGABAR")):
                    cur[w] += 1
    names = demangle(list(kernels))
    want = sys.argv[1:] or ["lsm_sweep_kernel<float, float, 3, false>", "lsm_sweep_kernel<float, float, 3, true>",
                            "lsm_sweep_kernel<double, double, 3, false>", "lsm_cluster_kernel<double, double, 3>",
                            "lsm_cluster_kernel<float, float, 3>", "lsm_step_tma_kernel<float, float, 3>",
                            "lsm_step_tma_kernel<float, double, 3>", "lsm_step_tma_kernel<double, double, 3>",
                            "lsm_step_tma_kernel<float, float, 8>", "lsm_solve_kernel<4>", "lsm_solve_kernel<9>",
                            "philox_quads_f32_kernel<10, true, 6>", "philox_quads_f32_kernel<7, true, 6>", "philox_paths_f64_kernel",
                            "normals_paths_kernel<double, true>", "normals_paths_kernel<float, true>", "sel_hist_kernel",
                            "first_hit_kernel<float>", "lean_walk_kernel<10, 0>"]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: static instruction counts per kernel (sm_100a)")
    print(f"# {len(kernels)} kernels in the library; tensor-core mnemonics (UTC*MMA / HMMA / LDTM) in any of them: "
          f"{sum(c['UTCHMMA'] + c['HMMA'] + c['LDTM'] for c in kernels.values())}")
    cols = [w for w in WATCH if w not in ("UTCHMMA", "HMMA", "LDTM", "UTMALDG")]
    print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{c[:6]:>6s}" for c in cols))
    for mangled, cnt in kernels.items():
        pretty = names.get(mangled, mangled)
        short = re.sub(r"\(.*", "", pretty).replace("amc::", "").replace("void ", "")
        if not any(w in short for w in want):
            continue
        print(f"{short[:58]:58s} {cnt['_total']:6d} " + " ".join(f"{cnt[c]:6d}" for c in cols))
    tot = collections.Counter()
    for cnt in kernels.values():
        tot.update(cnt)
    print(f"{'ALL KERNELS':58s} {tot['_total']:6d} " + " ".join(f"{tot[c]:6d}" for c in cols))


if __name__ == "__main__":
    main()
