# round 2, call P (2 GPUs): final tree -- GPU tier on one GPU's worth of tests plus the multi-GPU tests, smoke, and the
# driver's bench line at N=1 and N=2
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -12
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_default_g1.json 2> gpurun_out/r2p_default_g1.err; tail -2 gpurun_out/r2p_default_g1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2p_default_g2.json 2> gpurun_out/r2p_default_g2.err; tail -2 gpurun_out/r2p_default_g2.err
python - <<'PY'
import json
for g in (1, 2):
    d=json.load(open(f'gpurun_out/r2p_default_g{g}.json'))
    n=d['north_star_c3']
    print(g, 'c2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('traffic_frac_of_peak'), d.get('price_rel_err'), d.get('price_check',{}).get('within_4_se'), '| c3', n['value'], n['ms_per_step'], n['end_to_end_hbm']['frac_of_aggregate_copy_bandwidth'], n.get('price_matches_n1'), n['price_check']['within_4_se'])
PY
