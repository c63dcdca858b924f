#!/bin/bash
# parity tests + quick throughput probe of C2 (development loop)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/quick_bench.py 20 8192 32 1000 3 > gpurun_out/quick.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; cat gpurun_out/quick.log
