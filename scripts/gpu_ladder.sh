#!/bin/bash
# gpurun --gpus N: the GPU test suite (incl. the two-device test) and the NCCL ladder check on all N GPUs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    scripts/ladder_nccl_check.py > gpurun_out/ladder_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/ladder_nccl.log
tail -n 5 gpurun_out/pytest_gpu.log; grep -E "LADDER|rc=" gpurun_out/ladder_nccl.log
