# round 2, call N (1 GPU): the extended small-shape walk (path-free sets in every mode, the opt-in persistent sweep on
# stored sets) and the whole GPU tier once more
timeout 900 python -m pytest tests/test_gpu_small_shapes.py -q --tb=short 2>&1 | tail -12
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_default.json 2> gpurun_out/r2n_default.err; tail -2 gpurun_out/r2n_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_default.json')); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline'].get('traffic_frac_of_peak'), d['north_star_c3']['roofline']['traffic'], d['north_star_c3']['roofline'].get('traffic_frac_of_peak'))"
