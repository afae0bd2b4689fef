# end-of-round multi-GPU record on one 8-GPU box: parity test at 8, then c3 / c2 at N = 4 and 8, c4 at 8
timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -k p2p 2>&1 | tail -3
b() { # tag N workload steps warmup
  tag=$1; N=$2; wl=$3; st=$4; wu=$5
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps $st --warmup $wu > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err; tail -1 gpurun_out/${tag}.err | cut -c1-200
  python -c "
import json; d=json.load(open('gpurun_out/${tag}.json')); print('$tag', d['config'].get('allreduce'), '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price', d.get('price_grid_corners')), d['clocks'])"
}
b r1n_c3_g8 8 c3 5 2
b r1n_c3_g4 4 c3 5 2
b r1n_c2_g8 8 c2 10 3
b r1n_c2_g4 4 c2 10 3
b r1n_c4_g8 8 c4 2 1
b r1n_c5_g8 8 c5 3 1
