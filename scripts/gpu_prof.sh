#!/bin/bash
# one full ncu capture of the MH kernel in steady state (launch 13 = after DE joined the cycle)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mh_sorted_kernel -s 12 -c 1 -f -o gpurun_out/prof_mh \
    python scripts/quick_bench.py 20 8192 32 1000 1 > gpurun_out/prof_mh.log 2>&1
tail -n 3 gpurun_out/prof_mh.log
