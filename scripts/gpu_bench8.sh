#!/bin/bash
# gpurun --gpus N: NCCL ladder parity check, then bench.py at N GPUs in both sharding modes and the reference arm
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 scripts/ladder_nccl_check.py > gpurun_out/ladder_nccl_${NG}.log 2>&1; grep -E "LADDER|rror" gpurun_out/ladder_nccl_${NG}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_n${NG}.json 2> gpurun_out/bench_n${NG}.err; tail -n 3 gpurun_out/bench_n${NG}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 5 --warmup 3 --shard ladder > gpurun_out/bench_n${NG}_ladder.json 2> gpurun_out/bench_n${NG}_ladder.err; tail -n 5 gpurun_out/bench_n${NG}_ladder.err
if [ "$1" != "noref" ]; then timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $NG --steps 1 --warmup 1 > gpurun_out/bench_ref_n${NG}.json 2> gpurun_out/bench_ref_n${NG}.err; fi
cat gpurun_out/bench_n${NG}.json gpurun_out/bench_n${NG}_ladder.json
