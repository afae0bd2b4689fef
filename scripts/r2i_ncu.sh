# round 2, call I (call H's reports were larger than the 64 MiB that travel back): K1 grid A/B, launch lists of configs 2 and
# 3, 12 consecutive step-kernel launches with --cache-control none (a handful of metrics: small reports), and one
# --set full capture each of the step kernel (c2, 3 launches), K1 and K1z
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_$VAR
  timeout 600 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline --no-c3 "$@" > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err; tail -3 gpurun_out/r2i_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2i_$tag.json')); b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price'))"
}
VAR=k1b6; AMC_K1_BLOCKS=6 run c3 3 3
VAR=k1b8; AMC_K1_BLOCKS=8 run c3 3 3
VAR=k1b6again; AMC_K1_BLOCKS=6 run c3 3 3
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
C2="python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-c3"
C3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-c3"
$C2 > gpurun_out/r2i_c2_plain.json 2> gpurun_out/r2i_c2_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/r2i_launches_c2.csv $C2 > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none -k regex:lsm_step_tma -s 130 -c 12 --csv --log-file gpurun_out/r2i_step12_c2.csv $C2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsm_step_tma -s 130 -c 3 -o gpurun_out/r2i_step_c2 $C2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:normals_paths -s 4 -c 1 -o gpurun_out/r2i_k1z_c2 $C2 > /dev/null 2>&1
$C3 > gpurun_out/r2i_c3_plain.json 2> gpurun_out/r2i_c3_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1020 -c 520 --csv --log-file gpurun_out/r2i_launches_c3.csv $C3 > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none -k regex:lsm_step_tma -s 600 -c 12 --csv --log-file gpurun_out/r2i_step12_c3.csv $C3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:philox_quads -s 4 -c 1 -o gpurun_out/r2i_k1_c3 python bench.py --workload c3 --paths 25000000 --steps 1 --warmup 3 --no-cpu-baseline --no-c3 > /dev/null 2>&1
du -sh gpurun_out; ls -la gpurun_out
