#!/bin/bash
for c in 70 80 90; do echo "== sorted carveout $c"; PTMCMC_SORT_CARVEOUT=$c timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep "rep 1"; done
echo "== shadow (float queue)"; PTMCMC_MH_VARIANT=5 timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shadow" 2>&1 | tail -n 2
