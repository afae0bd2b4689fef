# round 2, call X8 (1 GPU): cluster kernel with asynchronous remote stores + per-CTA mbarrier instead of the per-pass
# cluster barrier -- tests, cluster arm of the A/B, trace
set -o pipefail
timeout 900 python -m pytest tests/test_gpu_small_shapes.py -x -q --tb=short 2>&1 | tail -12
AMC_CLUSTER=1 AMC_CLUSTER_MAX_PATHS=100000000 timeout 300 python scripts/r2x_cluster_ab.py > gpurun_out/r2x8_ab_cluster.jsonl 2> gpurun_out/r2x8_ab_cluster.err; tail -3 gpurun_out/r2x8_ab_cluster.err
python - <<'PY'
import json
for l in open('gpurun_out/r2x8_ab_cluster.jsonl'):
    x=json.loads(l)
    if x['kinds']==[2]: print(x['paths'], x['dtype'][-2:], x['basis'][:4], x['degree'], '%.3f ms (%.2f us/step)' % (x['median_ms'], x['us_per_step']))
PY
for cfg in "10000 Power 3" "100000 Power 3" "10000 Chebyshev 4"; do
  timeout 120 python scripts/r2x_cluster_trace.py $cfg 2>&1 | tail -2
done
