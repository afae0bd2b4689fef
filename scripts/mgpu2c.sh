timeout 500 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c2_g2_final.json 2> gpurun_out/c2_g2_final.err; tail -2 gpurun_out/c2_g2_final.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/c2_g2_final.json')); print('c2 N=2', d['config']['allreduce'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.4g %.2f ms'%(d['e2e']['value'], d['e2e']['ms_per_step']), {k:round(v,3) for k,v in d['breakdown_ms'].items()}, d['price'])"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5 N=2', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, d['price'])"
