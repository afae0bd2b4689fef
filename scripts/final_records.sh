run() { wl=$1; st=$2; wu=$3; shift 3
  timeout 300 python bench.py --workload $wl --steps $st --warmup $wu "$@" > gpurun_out/r1n_${wl}_g1.json 2> gpurun_out/r1n_${wl}_g1.err; tail -2 gpurun_out/r1n_${wl}_g1.err
  python -c "
import json; d=json.load(open('gpurun_out/r1n_${wl}_g1.json')); b=d['breakdown_ms']; print('$wl', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], d.get('price', d.get('price_grid_corners')), 'cpu %.3g'%d['cpu_baseline']['value'])"
}
run c3 3 1
run c4 2 1
run c5 3 1
run c1 20 3
