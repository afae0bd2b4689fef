# multi-GPU check: parity test under both transports, then c2 / c3 benches with the fused peer-memory all-reduce and NCCL
N=${1:-2}
timeout 400 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -15
for mode in p2p nccl; do
  for wlk in "c2 10 3" "c3 3 1"; do
    set -- $wlk
    AMC_ALLREDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $1 --steps $2 --warmup $3 > gpurun_out/${1}_g${N}_$mode.json 2> gpurun_out/${1}_g${N}_$mode.err; tail -3 gpurun_out/${1}_g${N}_$mode.err | cut -c1-300
    python -c "
import json; d=json.load(open('gpurun_out/${1}_g${N}_$mode.json')); print('$1 N=$N $mode', d['config']['allreduce'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d['price'])"
  done
done
