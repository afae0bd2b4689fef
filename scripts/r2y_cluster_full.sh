# round 2, call Y (1 GPU): whole GPU tier on the tree with the cluster kernel and the faster solve, then c1 through bench.py
# with and without the cluster kernel
set -o pipefail
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -12
timeout 300 python bench.py --workload c1 --steps 50 --warmup 10 > gpurun_out/r2y_c1_cluster.json 2> gpurun_out/r2y_c1_cluster.err; tail -2 gpurun_out/r2y_c1_cluster.err
AMC_CLUSTER=0 timeout 300 python bench.py --workload c1 --steps 50 --warmup 10 > gpurun_out/r2y_c1_chain.json 2> gpurun_out/r2y_c1_chain.err; tail -2 gpurun_out/r2y_c1_chain.err
python - <<'PY'
import json
for tag in ("cluster", "chain"):
    d=json.load(open(f'gpurun_out/r2y_c1_{tag}.json'))
    print(tag, d['value'], d['ms_per_step'], d['breakdown_ms'], d['roofline']['kernel'][:40], d.get('price_rel_err'), d['e2e']['ms_per_step'])
PY
