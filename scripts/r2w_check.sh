# round 2, call W (1 GPU): new test of the production solve route on the rank-truncated goldens + the random-contract rank
# exploration of round 1 (300 contracts, degree 4..10) on the warp SVD
timeout 600 python -m pytest tests/test_gpu_parity.py -q --tb=short -k "rank_truncated" 2>&1 | tail -8
timeout 900 python scripts/explore_rank.py 2>&1 | tail -12
