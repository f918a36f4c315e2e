#!/bin/bash
# A second build of the library with extra nvcc flags for ONE translation unit: build/alt/lib_<name>.so (same-box A/B: tools/ab_libs.sh,
# tools/ab_bench_libs.sh).  usage: build_alt.sh <name> "<flags>" [unit, default srcnn_tc2]
set -e
cd "$(dirname "$0")/.."
U=${3:-srcnn_tc2}
mkdir -p build/alt/obj_$1
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math --fmad=true -Wno-deprecated-gpu-targets"
$NV $2 -Xptxas -v -c srcnn_cpp_b200/csrc/$U.cu -o build/alt/obj_$1/$U.o 2>&1 | grep -E "Used|spill" | grep -v " 0 bytes spill" | head -8 || true
OBJS=$(ls build/obj/*.o | grep -v /$U.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -Xcompiler -fPIC -o build/alt/lib_$1.so $OBJS build/alt/obj_$1/$U.o -lnvjpeg_static -lculibos -lcuda
ls -la build/alt/lib_$1.so
