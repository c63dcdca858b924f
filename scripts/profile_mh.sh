mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mh_reg_kernel -s 12 -c 2 -f -o gpurun_out/prof_mh python scripts/quick_bench.py 20 8192 32 1000 1 > gpurun_out/prof_mh.log 2>&1
ls -la gpurun_out
