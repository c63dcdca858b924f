#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "register_kernel or curved or truncated_box" 2>&1 | tail -n 3
echo "== sorted (vector DE loads)"; timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep
for c in 60 100; do echo "== carveout $c"; PTMCMC_SORT_CARVEOUT=$c timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep; done
