#!/bin/bash
mkdir -p gpurun_out
( for nc in 96 128 64; do echo "== MMA C2 nc=$nc"; PTMCMC_MMA_NC=$nc PTMCMC_MH_VARIANT=3 timeout 300 python scripts/quick_bench.py 20 8192 32 1000 2; done ) 2>&1 | tee gpurun_out/mma_bench.log | grep -E "==|rep"
