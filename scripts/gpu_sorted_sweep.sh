#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "register_kernel or truncated or curved" 2>&1 | tail -n 2
for mb in 2 3; do echo "== PTMCMC_SORT_MINB=$mb"; PTMCMC_SORT_MINB=$mb timeout 300 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep; done
PTMCMC_SORT_MINB=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "register_kernel or truncated or curved" 2>&1 | tail -n 2
