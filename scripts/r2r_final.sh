# round 2, call R (1 GPU): last check of the tree as committed -- the whole GPU tier (the C example now prices a path-free set too)
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|Error|error|FAILED|^E " | tail -12
