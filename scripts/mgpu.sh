N=${1:-2}
timeout 200 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/c2_g$N.json 2> gpurun_out/c2_g$N.err; tail -3 gpurun_out/c2_g$N.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/c2_g$N.json')); print('c2 N=$N %.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d['price'], 'e2e %.4g'%d['e2e']['value'])"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c3 --steps 3 --warmup 1 > gpurun_out/c3_g$N.json 2> gpurun_out/c3_g$N.err; tail -3 gpurun_out/c3_g$N.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/c3_g$N.json')); print('c3 N=$N %.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d['price'])"
