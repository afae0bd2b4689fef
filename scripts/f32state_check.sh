timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweeps.py -q --tb=short -x -k "not slow" 2>&1 | tail -15
for st in float32 float64; do
for cfg in "c3 3 1" "c5 3 1" "c4 2 1"; do
  set -- $cfg
  timeout 300 python bench.py --workload $1 --state $st --steps $2 --warmup $3 --no-cpu-baseline > gpurun_out/$1_$st.json 2> gpurun_out/$1_$st.err; tail -3 gpurun_out/$1_$st.err
  python -c "
import json; d=json.load(open('gpurun_out/$1_$st.json')); print('$1 $st', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], 'launch us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d.get('price', d.get('price_grid_corners')))"
done; done
