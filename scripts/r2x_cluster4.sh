# round 2, call X4 (1 GPU): cluster kernel with the recursive-halving reduction, pairwise gather, ILP 8/4/2 -- tests, A/B, trace
set -o pipefail
timeout 900 python -m pytest tests/test_gpu_small_shapes.py -x -q --tb=short 2>&1 | tail -12
AMC_CLUSTER=1 AMC_CLUSTER_MAX_PATHS=100000000 timeout 300 python scripts/r2x_cluster_ab.py > gpurun_out/r2x4_ab_cluster.jsonl 2> gpurun_out/r2x4_ab_cluster.err; tail -3 gpurun_out/r2x4_ab_cluster.err
AMC_CLUSTER=0 timeout 300 python scripts/r2x_cluster_ab.py > gpurun_out/r2x4_ab_chain.jsonl 2> gpurun_out/r2x4_ab_chain.err; tail -3 gpurun_out/r2x4_ab_chain.err
python - <<'PY'
import json
a=[json.loads(l) for l in open('gpurun_out/r2x4_ab_cluster.jsonl')]
b=[json.loads(l) for l in open('gpurun_out/r2x4_ab_chain.jsonl')]
for x,y in zip(a,b):
    print(x['paths'], x['dtype'][-2:], x['basis'][:4], x['degree'], 'kinds', x['kinds'], y['kinds'], 'cluster %.3f ms (%.2f us/step)  chain %.3f ms (%.2f us/step)  price diff %.1e' % (x['median_ms'], x['us_per_step'], y['median_ms'], y['us_per_step'], abs(x['price']-y['price'])/y['price']))
PY
for cfg in "10000 Power 3" "100000 Power 3" "100000 Power 3 float32" "10000 Chebyshev 4"; do
  timeout 120 python scripts/r2x_cluster_trace.py $cfg 2>&1 | tail -2
done
