# round 2, call Z2 (2 GPUs): the multi-GPU tests and the N=2 bench line on the final tree (solve changes + cluster kernel)
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2z_default_g2.json 2> gpurun_out/r2z_default_g2.err; tail -2 gpurun_out/r2z_default_g2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_default_g2.json'))
n=d['north_star_c3']
print('N=2 c2', d['value'], d['ms_per_step'], d['roofline']['frac'], '| c3', n['value'], n['ms_per_step'], n.get('price_matches_n1'), n['price_check']['within_4_se'])
PY
