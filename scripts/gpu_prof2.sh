#!/bin/bash
mkdir -p gpurun_out
true \
    
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mh_mma_kernel -s 12 -c 1 -f -o gpurun_out/prof_mma_c3 \
    python scripts/config_bench.py c3 200 1 > gpurun_out/prof_mma_c3.log 2>&1
tail -n 2 gpurun_out/prof_mma_c2.log gpurun_out/prof_mma_c3.log
