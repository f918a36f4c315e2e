#!/bin/bash
# same-box A/B of several builds of the library: build/alt/lib_<name>.so, each timed by tools/ab_tc2.py, two rounds
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for round in 1 2; do
  for lib in build/alt/lib_*.so; do
    echo "== $lib (round $round)"
    SRCNN_B200_LIB=$PWD/$lib timeout 300 python tools/ab_tc2.py 2>&1 | grep -v CUDAEvent | cut -c1-200
  done
done | tee gpurun_out/ab_libs.log
