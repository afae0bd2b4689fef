# per-step fixed cost: small path counts where the streaming time is short
for cfg in "c1 100000 20" "c3 12500000 5" "c3 1000000 5" "c2 1000000 10"; do
  set -- $cfg
  timeout 200 python bench.py --workload $1 --paths $2 --steps $3 --warmup 3 --no-cpu-baseline > gpurun_out/lat_$1_$2.json 2> gpurun_out/lat_$1_$2.err; tail -2 gpurun_out/lat_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/lat_$1_$2.json')); n=d['config']['time_steps']; print('$1 P=$2', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['breakdown_ms'].items()}, 'per-step us: sweep %.2f step-kernel %.2f'%(1e3*d['breakdown_ms']['sweep_total']/(n+1), 1e3*d['roofline']['avg_launch_ms']), 'frac %.3f'%d['roofline']['frac'])"
done
