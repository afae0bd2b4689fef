"""Turn gpurun_out ncu artefacts into the small text/JSON summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/r1_launches_c2.csv  > profiles/r1_launches_c2.txt
    python profiles/summarize_ncu.py kernel   gpurun_out/r1_prof_step_c2.ncu-rep > profiles/r1_step_kernel_c2.txt
    python profiles/summarize_ncu.py steps    gpurun_out/r2i_step12_c2.csv > profiles/r2i_step12_c2.txt
    python profiles/summarize_ncu.py traffic  c2 10000000 51 gpurun_out/r2i_step12_c2.csv [c3 100000000 253 ...csv]
        -> profiles/step_kernel_traffic.json: DRAM bytes per sweep of the dominant kernel from an
           `ncu --set full --cache-control none` capture of >= 10 consecutive launches (mean per launch x launches per
           sweep); bench.py reads it into roofline.traffic
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
           "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__cycles_active.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':92s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:92s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {v[1] / tot:7.1%}")


def read_metric_csv(path):
    """`ncu --metrics ... --csv --log-file`: one row per (launch, metric) -> list of {metric: value in base units}."""
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
                "ms": 1e3, "msecond": 1e3}.get(r["Metric Unit"], 1.0)
        d = per.setdefault(r["ID"], {"kernel": r["Kernel Name"]})
        d[r["Metric Name"]] = v * mult
    return list(per.values())


def steps(path):
    """Table of consecutive step-kernel launches captured with --cache-control none."""
    rows = read_metric_csv(path)
    print(f"# {path}: ncu --metrics (below) --clock-control none --cache-control none, {len(rows)} consecutive launches of one sweep")
    print(f"# kernel: {rows[0]['kernel'][:100]}")
    print(f"{'launch':>6s} {'us':>8s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s} {'l2_hit_%':>9s} {'fp64_%':>7s} {'issue_%':>8s} {'warps_%':>8s}")
    for i, r in enumerate(rows):
        print(f"{i:6d} {r['gpu__time_duration.sum']:8.1f} {r['dram__bytes_read.sum'] / 1e6:11.1f} {r['dram__bytes_write.sum'] / 1e6:11.1f} "
              f"{r['lts__t_sector_hit_rate.pct']:9.1f} {r['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']:7.1f} "
              f"{r['smsp__issue_active.avg.pct_of_peak_sustained_active']:8.1f} {r['sm__warps_active.avg.pct_of_peak_sustained_active']:8.1f}")
    n = len(rows)
    rd = sum(r['dram__bytes_read.sum'] for r in rows) / n
    wr = sum(r['dram__bytes_write.sum'] for r in rows) / n
    us = sum(r['gpu__time_duration.sum'] for r in rows) / n
    print(f"# mean per launch: {us:.1f} us, DRAM read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB "
          f"-> {(rd + wr) / us / 1e6:.2f} TB/s under ncu")


def traffic(args):
    import json
    import os
    out = {}
    for i in range(0, len(args), 4):
        workload, paths, launches_per_sweep, rep = args[i], int(args[i + 1]), int(args[i + 2]), args[i + 3]
        rows = read_metric_csv(rep)
        n = len(rows)
        rd = sum(r["dram__bytes_read.sum"] for r in rows)
        wr = sum(r["dram__bytes_write.sum"] for r in rows)
        dur = sum(r["gpu__time_duration.sum"] for r in rows)
        out[workload] = {
            "kernel": rows[0]["kernel"][:80], "paths_per_gpu": paths, "launches_captured": n,
            "dram_read_per_launch": rd / n, "dram_write_per_launch": wr / n,
            "dram_bytes_per_launch": (rd + wr) / n, "launches_per_sweep": launches_per_sweep,
            "dram_bytes_per_sweep": (rd + wr) / n * launches_per_sweep,
            "mean_launch_us_under_ncu": dur / n,
            "source": f"ncu --metrics dram__bytes_read/write.sum --clock-control none --cache-control none, {n} consecutive "
                      f"launches of one sweep ({os.path.basename(rep)}; table in profiles/)"}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "step_kernel_traffic.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {path}: ncu --set full --clock-control none; one column per captured launch")
    print("kernels:", [r[name_i][:70] for r in data])
    for i, h in enumerate(hdr):
        if h in METRICS:
            print(f"{h:70s} {units[i]:14s} {[r[i] for r in data]}")
    # top warp-stall reasons of the first captured launch (share of warp-active cycles)
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_warp_active.pct") and data:
            try:
                stalls.append((float(data[0][i].replace(",", "")), h))
            except ValueError:
                pass
    for v, h in sorted(stalls, reverse=True)[:7]:
        print(f"{h:90s} {v:8.2f}")
    for i, h in enumerate(hdr):
        if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum") and data:
            try:
                v = float(data[0][i].replace(",", ""))
            except ValueError:
                continue
            if v > 0:
                print(f"{h:70s} {units[i]:14s} {data[0][i]}")


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        {"launches": launches, "kernel": kernel, "steps": steps}[sys.argv[1]](sys.argv[2])
