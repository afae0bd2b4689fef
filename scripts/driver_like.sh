set -x
time python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
time python bench.py > gpurun_out/driver_bench.json 2> gpurun_out/driver_bench.err; tail -3 gpurun_out/driver_bench.err; wc -l gpurun_out/driver_bench.json
python -c "
import json; d=json.load(open('gpurun_out/driver_bench.json')); print({k:d[k] for k in ['value','ms_per_step','n_gpus','steps','warmup','gpu_launches','dtype','scaling']}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks'])"
time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/driver_ref.json 2> gpurun_out/driver_ref.err; tail -3 gpurun_out/driver_ref.err; cat gpurun_out/driver_ref.json | cut -c1-700
