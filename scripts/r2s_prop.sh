# round 2, call S (1 GPU): the new property test of path-free sets
timeout 900 python -m pytest tests/test_gpu_property.py -q --tb=short 2>&1 | tail -15
