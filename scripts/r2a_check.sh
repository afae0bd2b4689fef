# round 2, call A: all GPU tests on the new float generator (quad counters, integer log-price) + K1 timing at config 3
mkdir -p gpurun_out
rm -f gpurun_out/big_shape_parity.jsonl
timeout 1700 python -m pytest tests -m gpu -q --tb=short -x -rP 2>&1 | grep -E "BIG_SHAPE_PARITY|passed|failed|Error|error|FAILED|assert" | tail -40
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')
  timeout 400 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline "$@" > gpurun_out/r2a_$tag.json 2> gpurun_out/r2a_$tag.err; tail -3 gpurun_out/r2a_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2a_$tag.json')); n=d['config']['time_steps']; b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], 'launch us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d.get('price', d.get('price_grid_corners')), d.get('price_rel_err'))"
}
run c3 3 3
AMC_PHILOX_ROUNDS=7 run c3 3 3 --paths 100000000
run c2 10 3
run c1 20 3
