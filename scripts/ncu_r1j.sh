# ncu evidence after the f32-state / batch work (each ncu command preceded by the same command run plain)
set -x
C3="python bench.py --workload c3 --paths 20000000 --steps 1 --warmup 1 --no-cpu-baseline"
C5="python bench.py --workload c5 --paths 20000000 --steps 1 --warmup 1 --no-cpu-baseline"
C4="python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline"
$C3 > gpurun_out/plain_c3j.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsm_step -s 300 -c 2 -o gpurun_out/r1j_step_c3_f32state $C3 > gpurun_out/ncu_j_c3.log 2>&1
$C5 > gpurun_out/plain_c5j.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsm_step -s 120 -c 2 -o gpurun_out/r1j_step_c5 $C5 > gpurun_out/ncu_j_c5.log 2>&1
$C4 > gpurun_out/plain_c4j.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsm_step -s 200 -c 2 -o gpurun_out/r1j_step_c4 $C4 > gpurun_out/ncu_j_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1j_launches_c4.csv $C4 > gpurun_out/ncu_l_c4.log 2>&1
ls -la gpurun_out | tail -8
