#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_srcnn_tc -s 3 -c 1 -f -o gpurun_out/prof_tc \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_bench.log 2>&1
tail -1 gpurun_out/prof_bench.log | cut -c1-80
