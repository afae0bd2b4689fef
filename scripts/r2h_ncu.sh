# round 2, call H: K1 grid A/B (one resident wave, 6 or 8 blocks per SM), then the ncu evidence of the round:
#   launch lists (gpu__time_duration) of configs 2 and 3, and --set full --cache-control none captures of >= 10 consecutive
#   step-kernel launches (DRAM traffic with the caches as the real pipeline leaves them), K1 and K1z
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_$VAR
  timeout 600 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline --no-c3 "$@" > gpurun_out/r2h_$tag.json 2> gpurun_out/r2h_$tag.err; tail -3 gpurun_out/r2h_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2h_$tag.json')); b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price'))"
}
VAR=k1b6; AMC_K1_BLOCKS=6 run c3 3 3
VAR=k1b8; AMC_K1_BLOCKS=8 run c3 3 3
VAR=k1b6; AMC_K1_BLOCKS=6 run c3 3 3 --paths 12500000
VAR=k1b8; AMC_K1_BLOCKS=8 run c3 3 3 --paths 12500000
VAR=k1b6r7; AMC_PHILOX_ROUNDS=7 AMC_K1_BLOCKS=6 run c3 3 3
C2="python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-c3"
C3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-c3"
$C2 > gpurun_out/r2h_c2_plain.json 2> gpurun_out/r2h_c2_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/r2h_launches_c2.csv $C2 > gpurun_out/r2h_launches_c2.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:lsm_step_tma -s 130 -c 12 -o gpurun_out/r2h_step_c2 $C2 > gpurun_out/r2h_step_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:normals_paths -s 4 -c 2 -o gpurun_out/r2h_k1z_c2 $C2 > gpurun_out/r2h_k1z_c2.log 2>&1
$C3 > gpurun_out/r2h_c3_plain.json 2> gpurun_out/r2h_c3_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1020 -c 520 --csv --log-file gpurun_out/r2h_launches_c3.csv $C3 > gpurun_out/r2h_launches_c3.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:lsm_step_tma -s 600 -c 12 -o gpurun_out/r2h_step_c3 $C3 > gpurun_out/r2h_step_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:philox_quads -s 4 -c 1 -o gpurun_out/r2h_k1_c3 $C3 > gpurun_out/r2h_k1_c3.log 2>&1
tail -n 2 gpurun_out/r2h_*.log
ls -la gpurun_out/r2h_*
