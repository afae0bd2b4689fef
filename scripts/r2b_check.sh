# round 2, call B: persistent one-launch sweep + path-free sets: new tests first (every spin has an 8 s limit), then all
mkdir -p gpurun_out
rm -f gpurun_out/big_shape_parity.jsonl
timeout 600 python -m pytest tests/test_gpu_lean.py tests/test_gpu_generator.py -q --tb=short -x 2>&1 | tail -15
timeout 1700 python -m pytest tests -m gpu -q --tb=short -rP 2>&1 | grep -E "BIG_SHAPE_PARITY|FP32_SCALING|passed|failed|Error|error|FAILED|assert|^E " | tail -60
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')
  timeout 400 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline "$@" > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err; tail -3 gpurun_out/r2b_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2b_$tag.json')); n=d['config']['time_steps']; b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], 'launch us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d.get('price', d.get('price_grid_corners')), d.get('price_rel_err'))"
}
run c3 3 3
AMC_PREFILTER=0 run c3 3 3 --paths 100000000
run c3 3 3 --paths 12500000
AMC_PERSISTENT=0 run c3 3 3 --paths 12500000
run c2 10 3
AMC_PERSISTENT=0 run c2 10 3 --paths 10000000
run c1 20 3
run c5 3 3
