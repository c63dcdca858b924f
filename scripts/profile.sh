#!/bin/bash
# Run on the GPU box (under gpurun): launch list of the bench command + one full capture of the MH kernel.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mh_reg_kernel -s 2 -c 2 -f -o gpurun_out/prof_mh \
    python scripts/quick_bench.py 20 8192 32 200 1 > gpurun_out/prof_mh.log 2>&1
ls -la gpurun_out
