timeout 500 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -12
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "philox" 2>&1 | tail -3
for cfg in "c3 3 1" "c2 10 3"; do set -- $cfg
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $1 --steps $2 --warmup $3 > gpurun_out/$1_g2.json 2> gpurun_out/$1_g2.err; tail -2 gpurun_out/$1_g2.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/$1_g2.json')); print('$1 N=2', d['config']['allreduce'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d['price'])"
done
timeout 200 python bench.py --workload c3 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 N=1', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, d['price'])"
