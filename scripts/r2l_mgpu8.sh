# round 2, call L (8 GPUs): the driver's bench line at N=8 (c2 weak + north_star_c3 strong, price_matches_n1), the same at
# N=4, path-free c3 at N=8
PORT=29721
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2l_default_g$N.json 2> gpurun_out/r2l_default_g$N.err; tail -2 gpurun_out/r2l_default_g$N.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2l_default_g$N.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches','price']}, d.get('price_check',{}).get('within_4_se'), 'frac', d['roofline']['frac'], d['config']['allreduce'], 'e2e', d['e2e']['value'])
n=d['north_star_c3']; print('c3', {k:n[k] for k in ['value','ms_per_step','steps','price','breakdown_ms']}, n['end_to_end_hbm']['frac_of_aggregate_copy_bandwidth'], n.get('price_check',{}).get('within_4_se'), n.get('price_matches_n1'))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29750 bench.py --gpus 8 --workload c3 --lean --steps 3 --warmup 3 --no-c3 > gpurun_out/r2l_c3lean_g8.json 2> gpurun_out/r2l_c3lean_g8.err; tail -2 gpurun_out/r2l_c3lean_g8.err
python -c "
import json; d=json.load(open('gpurun_out/r2l_c3lean_g8.json')); print('c3lean g8', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], d['breakdown_ms'], d['price'], d.get('price_matches_n1'))"
