# round 2, call K (2 GPUs): the multi-GPU parity tests again after the rank-local column statistics fix
export AMC_SWEEP_DEBUG=1
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | grep -v "^E    *$" | tail -40
