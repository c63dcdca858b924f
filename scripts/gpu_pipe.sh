#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warp_specialised" 2>&1 | tail -n 12
for npw in 4 6 8; do echo "== PIPE NPW=$npw"; PTMCMC_PIPE_NPW=$npw PTMCMC_MH_VARIANT=4 timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep; done
