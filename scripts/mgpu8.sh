# 8-GPU confirmation: parity test (peer-memory transport), then c3 / c2 / c4 benches
N=${1:-8}
AMC_TEST_TRANSPORTS=p2p timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -k p2p 2>&1 | tail -5
b() { # tag env workload steps warmup
  tag=$1; mode=$2; wl=$3; st=$4; wu=$5
  AMC_ALLREDUCE=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps $st --warmup $wu > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err; tail -2 gpurun_out/${tag}.err | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/${tag}.json')); print('$tag', d['config'].get('allreduce'), '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price', d.get('price_grid_corners')))"
}
b c3_g${N}_p2p p2p c3 5 2
b c3_g${N}_nccl nccl c3 5 2
b c2_g${N}_p2p p2p c2 10 3
b c4_g${N} p2p c4 2 1
