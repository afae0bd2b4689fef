"""Turn gpurun_out ncu artefacts into the small text/JSON summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/r1_launches_c2.csv  > profiles/r1_launches_c2.txt
    python profiles/summarize_ncu.py kernel   gpurun_out/r1_prof_step_c2.ncu-rep > profiles/r1_step_kernel_c2.txt
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
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
