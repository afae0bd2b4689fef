# round 2, call T (1 GPU): float step kernels above degree 3 compiled for 4 (degree 4-5) / 3 (degree 6-10) resident blocks
# per SM instead of the 2 their unconstrained register allocation allows: A/B against the old build
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_$VAR
  timeout 300 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline --no-c3 "$@" > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err; tail -2 gpurun_out/r2t_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2t_$tag.json')); b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'frac %.3f'%d['roofline']['frac'], d['roofline']['bound'], d.get('price'))"
}
OLD=$PWD/american_monte_carlo_b200/libamc_oldblocks.so
for D in 4 5 6 7; do
  VAR=new; run c3 3 3 --degree $D --paths 50000000
  VAR=old; AMC_LIBAMC=$OLD run c3 3 3 --degree $D --paths 50000000
done
VAR=new; run c3 3 3 --degree 5 --paths 50000000 --state float64
VAR=old; AMC_LIBAMC=$OLD run c3 3 3 --degree 5 --paths 50000000 --state float64
