#!/bin/bash
# ncu captures (run under gpurun, 1 GPU): launch list of a short bench run + one full capture of the
# fused SRCNN kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_srcnn_tc -s 3 -c 1 -f -o gpurun_out/prof_tc \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/prof_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_color_bicubic -s 3 -c 1 -f -o gpurun_out/prof_a \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/prof_bench.log 2>&1
tail -5 gpurun_out/prof_bench.log
python bench.py --steps 50 --warmup 5 --no-cpu
