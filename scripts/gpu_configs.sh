#!/bin/bash
mkdir -p gpurun_out
( true
  timeout 900 python scripts/config_bench.py c3 200 2 ) 2>&1 | tee gpurun_out/configs.log
