# round 2, call D: the single cooperative sweep kernel (last block of a pass solves): tests, then persistent vs launch chain
export AMC_SWEEP_DEBUG=1
timeout 600 python -m pytest tests/test_gpu_lean.py tests/test_gpu_generator.py tests/test_gpu_small_shapes.py -q --tb=line -x 2>&1 | tail -8
timeout 1700 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -E "passed|failed|Error|error|FAILED|debug" | tail -30
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_$VAR
  timeout 400 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline "$@" > gpurun_out/r2d_$tag.json 2> gpurun_out/r2d_$tag.err; tail -3 gpurun_out/r2d_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2d_$tag.json')); n=d['config']['time_steps']; b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price', d.get('price_grid_corners')), d.get('price_rel_err'))"
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
VAR=nopre; export AMC_PREFILTER=0
run c3 3 3
run c3 3 3 --paths 12500000
