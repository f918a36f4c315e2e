#!/bin/bash
# one full ncu capture of the fused kernel with source-level counters; the report comes back in gpurun_out/
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_srcnn_tc2 -s 3 -c 1 -f -o gpurun_out/prof_tc2_src \
    python tools/tc2_one.py > gpurun_out/ncu_src.log 2>&1
tail -5 gpurun_out/ncu_src.log
ls -la gpurun_out/*.ncu-rep
