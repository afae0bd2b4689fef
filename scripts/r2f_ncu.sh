# round 2, call F: why is the cooperative sweep kernel slower than the launch chain on float columns?  ncu --set full of
# both on the 12.5M-path shard of config 3 (the plain run of each command line precedes its capture)
export AMC_SWEEP_DEBUG=1
CMD="python bench.py --workload c3 --paths 12500000 --steps 1 --warmup 3 --no-cpu-baseline --no-c3"
$CMD > gpurun_out/r2f_persist_plain.json 2> gpurun_out/r2f_persist_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:lsm_sweep_kernel -s 3 -c 1 -o gpurun_out/r2f_persist $CMD > gpurun_out/r2f_persist_ncu.log 2>&1
tail -2 gpurun_out/r2f_persist_plain.err
AMC_PERSISTENT=0 $CMD > gpurun_out/r2f_chain_plain.json 2> gpurun_out/r2f_chain_plain.err &&
AMC_PERSISTENT=0 ncu --set full --clock-control none --import-source on -k regex:lsm_step_tma -s 900 -c 3 -o gpurun_out/r2f_chain $CMD > gpurun_out/r2f_chain_ncu.log 2>&1
# K1 pipes
ncu --set full --clock-control none --import-source on -k regex:philox_quads -s 3 -c 1 -o gpurun_out/r2f_k1 $CMD > gpurun_out/r2f_k1_ncu.log 2>&1
tail -3 gpurun_out/r2f_persist_ncu.log gpurun_out/r2f_chain_ncu.log gpurun_out/r2f_k1_ncu.log
ls -la gpurun_out/*.ncu-rep
