# ncu evidence, end of round 1 (each ncu command preceded by the same command run plain)
set -x
export AMC_GRAPH=0     # per-kernel profiling: plain launches (a replayed graph is profiled as the same kernels anyway)
C2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
C3="python bench.py --workload c3 --paths 20000000 --steps 1 --warmup 1 --no-cpu-baseline"
$C2 > gpurun_out/plain_c2m.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1m_launches_c2.csv $C2 > gpurun_out/ncu_l_c2m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsm_step -s 60 -c 2 -o gpurun_out/r1m_step_c2 $C2 > gpurun_out/ncu_s_c2m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:normals_paths -s 2 -c 1 -o gpurun_out/r1m_k1z_c2 $C2 > gpurun_out/ncu_z_c2m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsm_solve -s 60 -c 2 -o gpurun_out/r1m_solve_c2 $C2 > gpurun_out/ncu_v_c2m.log 2>&1
$C3 > gpurun_out/plain_c3m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:philox_paths -s 1 -c 1 -o gpurun_out/r1m_philox_c3 $C3 > gpurun_out/ncu_p_c3m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsm_step -s 300 -c 2 -o gpurun_out/r1m_step_c3 $C3 > gpurun_out/ncu_s_c3m.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1m_launches_c3.csv $C3 > gpurun_out/ncu_l_c3m.log 2>&1
ls -la gpurun_out | grep r1m
