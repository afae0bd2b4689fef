# round 2, call J (2 GPUs): the multi-GPU parity tests under both transports (sharded stored and path-free sweeps, the
# rank-local ndarray case), then the driver's bench line at N=2 (c2 weak + north_star_c3 strong) and lean c3 at N=2
export AMC_SWEEP_DEBUG=1
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -15
PORT=29711
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_default_g2.json 2> gpurun_out/r2j_default_g2.err; tail -3 gpurun_out/r2j_default_g2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_default_g2.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches','price']}, d.get('price_check'))
print('roofline', d['roofline']['frac'], 'alg', d['config']['allreduce'])
n=d['north_star_c3']; print('c3', {k:n[k] for k in ['value','ms_per_step','steps','price','breakdown_ms','end_to_end_hbm']}, n.get('price_check'), n.get('price_matches_n1'), n.get('price_n1'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --workload c3 --lean --steps 3 --warmup 3 --no-c3 > gpurun_out/r2j_c3lean_g2.json 2> gpurun_out/r2j_c3lean_g2.err; tail -3 gpurun_out/r2j_c3lean_g2.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_c3lean_g2.json')); print('c3lean g2', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], d['breakdown_ms'], d['price'], d.get('price_matches_n1'))"
AMC_ALLREDUCE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --workload c3 --steps 3 --warmup 3 --no-c3 > gpurun_out/r2j_c3nccl_g2.json 2> gpurun_out/r2j_c3nccl_g2.err; tail -3 gpurun_out/r2j_c3nccl_g2.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_c3nccl_g2.json')); print('c3 nccl g2', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], d['breakdown_ms'], d['price'], d.get('price_matches_n1'), d['config']['allreduce'])"
