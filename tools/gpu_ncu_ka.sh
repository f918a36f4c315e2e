#!/bin/bash
# one ncu --set full capture of the colour+bicubic kernel (run under gpurun); report lands in gpurun_out/prof_ka_int.ncu-rep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_color_bicubic -s 6 -c 1 -f -o gpurun_out/prof_ka_int \
  python tools/ab_stages.py ${KA_ARGS:-} > gpurun_out/prof_ka_int.log 2>&1
tail -3 gpurun_out/prof_ka_int.log
