timeout 900 python -m pytest tests/test_gpu_ccr.py tests/test_gpu_parity.py tests/test_gpu_sweeps.py -q --tb=short -x -k "not slow" 2>&1 | tail -12
timeout 300 python scripts/sanitize_small.py 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 60 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -4 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; tail -4 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool synccheck python scripts/sanitize_small.py > gpurun_out/sanitize_synccheck.log 2>&1; tail -4 gpurun_out/sanitize_synccheck.log
