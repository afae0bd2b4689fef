timeout 300 python -m pytest tests -m "gpu and not slow" -q --tb=short -x 2>&1 | grep -v "^  " | tail -6
for v in 0 1; do
AMC_PDL=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/pdl_$v.json 2>gpurun_out/pdl_$v.err
python -c "
import json; d=json.load(open('gpurun_out/pdl_$v.json')); print('pdl=$v', '%.4g'%d['value'], '%.3f'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d['price_rel_err'])"
done
