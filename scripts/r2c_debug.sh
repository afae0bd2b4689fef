# round 2, call C: does force-loading the sweep kernels cure the first-use timeouts of call B?
export AMC_SWEEP_DEBUG=1
timeout 600 python -m pytest tests/test_gpu_lean.py tests/test_gpu_generator.py -q --tb=line -x 2>&1 | tail -15
timeout 1700 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -E "passed|failed|Error|error|FAILED|debug" | tail -30
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')
  timeout 400 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline "$@" > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err; tail -3 gpurun_out/r2c_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c_$tag.json')); n=d['config']['time_steps']; b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], 'launch us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d.get('price', d.get('price_grid_corners')), d.get('price_rel_err'))"
}
run c1 20 3
run c2 10 3
run c3 3 3
AMC_PREFILTER=0 run c3 3 3 --paths 100000000
run c3 3 3 --paths 12500000
run c5 3 3
