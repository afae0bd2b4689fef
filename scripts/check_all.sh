timeout 400 python -m pytest tests -m "gpu and not slow" -q --tb=short -x 2>&1 | grep -v "^  " | tail -6
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2.json 2>gpurun_out/c2.err; tail -2 gpurun_out/c2.err
python -c "
import json; d=json.load(open('gpurun_out/c2.json')); print('c2 %.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], 'launch_us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d['price_rel_err'])"
timeout 300 python bench.py --workload c3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c3.json 2>gpurun_out/c3.err; tail -2 gpurun_out/c3.err
python -c "
import json; d=json.load(open('gpurun_out/c3.json')); print('c3 %.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], 'launch_us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d['price'], d['clocks'])"
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c5.json 2>gpurun_out/c5.err; tail -2 gpurun_out/c5.err
python -c "
import json; d=json.load(open('gpurun_out/c5.json')); print('c5 %.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], 'launch_us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d['price'])"
