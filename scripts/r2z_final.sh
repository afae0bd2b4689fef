# round 2, call Z (1 GPU): final tree -- GPU tier, smoke, the driver's bench line, the reference arm (short)
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -12
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_default_g1.json 2> gpurun_out/r2z_default_g1.err; tail -2 gpurun_out/r2z_default_g1.err
timeout 300 python bench.py --workload c1 --steps 50 --warmup 10 --no-c3 > gpurun_out/r2z_c1.json 2> gpurun_out/r2z_c1.err; tail -2 gpurun_out/r2z_c1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_default_g1.json'))
n=d['north_star_c3']
print('c2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('traffic_frac_of_peak'), d.get('price_rel_err'), 'e2e', d['e2e']['ms_per_step'], '| c3', n['value'], n['ms_per_step'], n['end_to_end_hbm']['frac_of_aggregate_copy_bandwidth'], n.get('price_matches_n1'), n['price_check']['within_4_se'], d['clocks'])
c=json.load(open('gpurun_out/r2z_c1.json'))
print('c1', c['value'], c['ms_per_step'], c['breakdown_ms'], c['roofline']['kernel'][:30], c['gpu_launches'])
PY
