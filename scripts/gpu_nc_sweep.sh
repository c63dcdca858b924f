#!/bin/bash
mkdir -p gpurun_out
for nc in 128 160 192 224 256; do
  echo "== PTMCMC_SORT_NC=$nc"
  PTMCMC_SORT_NC=$nc timeout 300 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep
done | tee gpurun_out/nc_sweep.log
