timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')
  timeout 300 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline "$@" > gpurun_out/p_$tag.json 2> gpurun_out/p_$tag.err; tail -3 gpurun_out/p_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/p_$tag.json')); n=d['config']['time_steps']; b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], 'launch us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d.get('price', d.get('price_grid_corners')), d.get('price_rel_err'))"
}
run c5 3 1
run c3 3 1
run c2 10 3
