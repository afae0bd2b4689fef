# round 2, call G: launch chain as the default for stored sets, persistent kernel for path-free sets: all tests, the
# driver's bench lines, the path-free benches
export AMC_SWEEP_DEBUG=1
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|debug|^E " | tail -20
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_default.json 2> gpurun_out/r2g_default.err; tail -3 gpurun_out/r2g_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_default.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','price','price_rel_err']})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('clocks', d['clocks']); print('cpu', d.get('cpu_baseline'))
n=d['north_star_c3']; print('c3', {k:n[k] for k in ['value','ms_per_step','steps','price','breakdown_ms','end_to_end_hbm']}, n.get('price_check'), n['roofline']['frac'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err; cut -c1-900 gpurun_out/r2g_ref.json
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')
  timeout 600 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline --no-c3 "$@" > gpurun_out/r2g_$tag.json 2> gpurun_out/r2g_$tag.err; tail -3 gpurun_out/r2g_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2g_$tag.json')); b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price'), d.get('price_check',{}).get('within_4_se'))"
}
run c3 3 3 --lean
run c3 3 3 --lean --paths 12500000
run c3 3 3 --state float64
run c3 2 3 --lean --scaling --paths 1000000000
run c3 3 3 --scaling
run c1 20 3
run c5 3 3
run c4 2 3
