timeout 300 python -m pytest tests -m "gpu and not slow" -q --tb=short -x 2>&1 | grep -v "^  " | tail -6
for cfg in "0 0" "1 0" "0 1" "1 1"; do set -- $cfg
AMC_L2_REVERSE=$1 AMC_L2_HINTS=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$1$2.json 2>gpurun_out/ab_$1$2.err
python -c "
import json,sys; d=json.load(open('gpurun_out/ab_$1$2.json')); print('rev=$1 hints=$2', '%.4g'%d['value'], '%.3f'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], 'launch_us %.1f'%(1e3*d['roofline']['avg_launch_ms']), d['price_rel_err'])"
done
