timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweeps.py tests/test_gpu_api_parity.py tests/test_gpu_ccr.py -q --tb=short -x -k "not slow" 2>&1 | tail -5
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_g$AMC_GRAPH
  timeout 300 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline "$@" > gpurun_out/p_$tag.json 2> gpurun_out/p_$tag.err; tail -3 gpurun_out/p_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/p_$tag.json')); n=d['config']['time_steps']; b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'sweep/step us %.2f'%(1e3*b.get('sweep_total',0)/(n+1)), d.get('price', d.get('price_grid_corners')))"
}
for g in 1 0; do export AMC_GRAPH=$g
run c1 20 3
run c2 10 3 --paths 1000000
run c4 2 1
run c2 10 3
done
