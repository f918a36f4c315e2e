#!/bin/bash
# A second build of the library with extra nvcc flags for srcnn_tc2.cu: build/alt/lib_<name>.so (same-box A/B: tools/ab_libs.sh)
# usage: build_alt.sh <name> "<flags>"
set -e
cd "$(dirname "$0")/.."
mkdir -p build/alt/obj_$1
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math --fmad=true -Wno-deprecated-gpu-targets"
$NV $2 -c srcnn_cpp_b200/csrc/srcnn_tc2.cu -o build/alt/obj_$1/srcnn_tc2.o
OBJS=$(ls build/obj/*.o | grep -v srcnn_tc2.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -Xcompiler -fPIC -o build/alt/lib_$1.so $OBJS build/alt/obj_$1/srcnn_tc2.o -lnvjpeg_static -lculibos -lcuda
ls -la build/alt/lib_$1.so
