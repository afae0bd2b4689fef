# round 2, call V (1 GPU): warp-parallel Jacobi SVD for the rank-truncated steps: GPU tier, then sweep latency per
# configuration with the default solve selection
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -12
echo "--- default (warp routine from degree 6)"; python scripts/r2u_solve_latency.py 2>&1 | tail -16
