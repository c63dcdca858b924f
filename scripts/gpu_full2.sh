#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/config_bench.py c2 1000 3 2>&1 | grep -E "rep|class"
timeout 600 python scripts/config_bench.py c3 1000 2 2>&1 | grep -E "rep|class"
