# round 2, call O (1 GPU): the latency variant of the persistent kernel for small stored sets: all GPU tests (every golden
# case with <= 160k paths now runs on it), then sweep time per step against the launch chain at 25k .. 800k paths
export AMC_SWEEP_DEBUG=1
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E |debug" | tail -12
run() { wl=$1; st=$2; wu=$3; shift 3; tag=$wl$(echo "$*" | tr -d ' -')_$VAR
  timeout 300 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline --no-c3 "$@" > gpurun_out/r2o_$tag.json 2> gpurun_out/r2o_$tag.err; tail -2 gpurun_out/r2o_$tag.err | grep -v debug
  python -c "
import json; d=json.load(open('gpurun_out/r2o_$tag.json')); b=d['breakdown_ms']; print('$tag', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, d.get('price'))"
}
for P in 25000 100000 200000 400000 800000; do
  VAR=chain; AMC_SMALL_PATHS=0 run c1 20 3 --paths $P
  VAR=small; AMC_SMALL_PATHS=100000000 run c1 20 3 --paths $P
done
VAR=chain; AMC_SMALL_PATHS=0 run c3 10 3 --paths 100000
VAR=small; AMC_SMALL_PATHS=100000000 run c3 10 3 --paths 100000
