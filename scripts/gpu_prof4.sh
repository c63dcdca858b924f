#!/bin/bash
mkdir -p gpurun_out
PTMCMC_PIPE_NPW=6 PTMCMC_MH_VARIANT=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mh_pipe_kernel -s 12 -c 1 -f -o gpurun_out/prof_pipe \
    python scripts/quick_bench.py 20 8192 32 1000 1 > gpurun_out/prof_pipe.log 2>&1
tail -n 2 gpurun_out/prof_pipe.log
