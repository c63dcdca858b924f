#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shadow" 2>&1 | tail -n 12
echo "== sorted"; timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep
echo "== shadow"; PTMCMC_MH_VARIANT=5 timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep
