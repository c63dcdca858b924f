#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
echo "== sorted"; timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep -E "rep|accept"
echo "== pipe6"; PTMCMC_MH_VARIANT=4 timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep
echo "== mma"; PTMCMC_MH_VARIANT=3 timeout 120 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep rep
timeout 300 python scripts/config_bench.py c3 500 2 2>&1 | grep -E "rep"
timeout 300 python scripts/config_bench.py c4 1000 2 2>&1 | grep -E "rep"
