# round 2, call E: cooperative sweep kernel after the register clean-up (no prefilter, solve out of line, grid-constant
# parameters): persistent vs launch chain, every workload
export AMC_SWEEP_DEBUG=1
timeout 900 python -m pytest tests -m gpu -q --tb=line -x 2>&1 | grep -E "passed|failed|Error|error|FAILED|debug" | tail -10
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_$VAR
  timeout 400 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline --no-c3 "$@" > gpurun_out/r2e_$tag.json 2> gpurun_out/r2e_$tag.err; tail -3 gpurun_out/r2e_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2e_$tag.json')); b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price'), d.get('price_rel_err'), d.get('price_check',{}).get('within_4_se'))"
}
for VAR in persist chain; do
  if [ $VAR = chain ]; then export AMC_PERSISTENT=0; fi
  run c1 20 3
  run c2 10 3
  run c3 3 3
  run c3 3 3 --paths 12500000
  run c5 3 3
done
unset AMC_PERSISTENT
VAR=lean
run c3 3 3 --lean
run c3 3 3 --lean --paths 12500000
run c3 2 3 --lean --paths 1000000000
