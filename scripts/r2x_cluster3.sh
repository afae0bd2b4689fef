# round 2, call X3 (1 GPU): phase trace of the cluster kernel
for cfg in "10000 Power 3" "100000 Power 3" "100000 Power 3 float32" "10000 Chebyshev 4" "10000 Power 5" "10000 Power 1"; do
  timeout 120 python scripts/r2x_cluster_trace.py $cfg 2>&1 | tail -2
done
