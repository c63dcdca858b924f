#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core" > gpurun_out/pytest_mma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mma.log
tail -n 25 gpurun_out/pytest_mma.log
( PTMCMC_MH_VARIANT=3 timeout 300 python scripts/quick_bench.py 20 8192 32 1000 2
  timeout 600 python scripts/config_bench.py c3 500 2 ) 2>&1 | tee gpurun_out/mma_bench.log | grep -E "rep|class|accept|var"
