#!/bin/bash
mkdir -p gpurun_out
PTMCMC_MMA_NC=96 PTMCMC_MH_VARIANT=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mh_mma_kernel -s 12 -c 1 -f -o gpurun_out/prof_mma_c2 \
    python scripts/quick_bench.py 20 8192 32 1000 1 > gpurun_out/prof_mma_c2.log 2>&1
tail -n 2 gpurun_out/prof_mma_c2.log
