#!/bin/bash
# Round evidence on one B200 (under gpurun): parity tests, smoke, bench (both arms), ncu launch list of the
# bench command and full captures of the dominant kernels.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-side > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mh_sorted_kernel -s 12 -c 1 -f -o gpurun_out/prof_mh_r02 \
    python scripts/quick_bench.py 20 8192 32 1000 1 > gpurun_out/prof_mh_r02.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mh_mma_split_kernel -s 12 -c 1 -f -o gpurun_out/prof_mma_c3_r02 \
    python scripts/config_bench.py c3 200 1 > gpurun_out/prof_mma_c3_r02.log 2>&1
ls -la gpurun_out
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log
tail -c 600 gpurun_out/bench_ref.json
