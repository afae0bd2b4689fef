run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_w$AMC_WARP_SOLVE
  timeout 300 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline "$@" > gpurun_out/p_$tag.json 2> gpurun_out/p_$tag.err; tail -3 gpurun_out/p_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/p_$tag.json')); n=d['config']['time_steps']; b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'sweep/step us %.2f'%(1e3*b.get('sweep_total',0)/(n+1)), 'solve us %.2f'%(1e3*b.get('solve_kernels',0)/(n+1)), d.get('price'))"
}
for w in 0 1; do export AMC_WARP_SOLVE=$w
run c5 10 3 --paths 1000000
run c1 20 3
done
