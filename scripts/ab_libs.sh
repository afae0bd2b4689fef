for v in A B A B; do
AMC_LIBAMC=$PWD/american_monte_carlo_b200/libamc_$v.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(x,3) for k,x in d['breakdown_ms'].items()}, 'frac %.3f'%d['roofline']['frac'], d['price_rel_err'])"
done
