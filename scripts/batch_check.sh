timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweeps.py -q --tb=short -x -k "batch or sweeps or grid" 2>&1 | tail -15
timeout 300 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c4.json 2> gpurun_out/c4.err; tail -5 gpurun_out/c4.err
python -c "
import json; d=json.load(open('gpurun_out/c4.json')); print('c4 %.4g'%d['value'], '%.3f ms'%d['ms_per_step'], d['breakdown_ms'], 'frac %.3f'%d['roofline']['frac'], 'launch us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d['price_grid_corners'], d['clocks'])"
