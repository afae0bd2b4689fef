# round 2, call Q (8 GPUs): the driver's bench line at N=8 and N=4 on the final tree
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29760+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2q_default_g$N.json 2> gpurun_out/r2q_default_g$N.err; tail -2 gpurun_out/r2q_default_g$N.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2q_default_g$N.json'))
n=d['north_star_c3']
print($N, 'c2', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('price_check',{}).get('within_4_se'), 'e2e', d['e2e']['value'], '| c3', n['value'], n['ms_per_step'], n['breakdown_ms'], n['end_to_end_hbm']['frac_of_aggregate_copy_bandwidth'], n.get('price_matches_n1'), n['price_check']['within_4_se'])
PY
done
