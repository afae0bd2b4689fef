# round 2, call M (1 GPU): the tree as it stands -- all GPU tests, smoke, the driver's two bench lines; then two A/Bs:
# path-free sweep at 3 resident blocks per SM (variant build), and the chunk size of the host -> device normals pipeline
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -12
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_default.json 2> gpurun_out/r2m_default.err; tail -3 gpurun_out/r2m_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_default.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','price','price_rel_err']}, d['breakdown_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('roofline', {k:d['roofline'][k] for k in ['frac','achieved','traffic']}); print('clocks', d['clocks'])
n=d['north_star_c3']; print('c3', {k:n[k] for k in ['value','ms_per_step','steps','price','breakdown_ms']}, n['end_to_end_hbm']['frac_of_aggregate_copy_bandwidth'], n.get('price_check',{}).get('within_4_se'), n.get('price_matches_n1'), n['roofline']['traffic'])
PY
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; cut -c1-700 gpurun_out/r2m_ref.json
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_$VAR
  timeout 600 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline --no-c3 "$@" > gpurun_out/r2m_$tag.json 2> gpurun_out/r2m_$tag.err; tail -3 gpurun_out/r2m_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2m_$tag.json')); b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'e2e ms %.2f'%d['e2e']['ms_per_step'], d.get('price'))"
}
VAR=lean2; run c3 3 3 --lean
VAR=lean3; AMC_LIBAMC=$PWD/american_monte_carlo_b200/libamc_lean3.so run c3 3 3 --lean
VAR=lean3s; AMC_LIBAMC=$PWD/american_monte_carlo_b200/libamc_lean3.so run c3 3 3 --lean --paths 12500000
VAR=lean2s; run c3 3 3 --lean --paths 12500000
for MB in 64 256 1024; do VAR=chunk$MB; AMC_STAGE_CHUNK_MB=$MB run c2 10 3; done
