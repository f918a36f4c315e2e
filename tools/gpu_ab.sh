#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for flags in "-DSRCNN_NO_SETMAXNREG" ""; do
  echo "##### variant: [$flags]"
  rm -f build/obj/srcnn_tc.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math $flags -c srcnn_cpp_b200/csrc/srcnn_tc.cu -o build/obj/srcnn_tc.o 2>&1 | grep -i error
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -Xcompiler -fPIC -o srcnn_cpp_b200/libsrcnn_b200.so build/obj/api.o build/obj/color_bicubic.o build/obj/srcnn_fp32.o build/obj/srcnn_tc.o build/obj/weights_blob.o -lcuda
  timeout 300 python -m pytest tests/test_stage_parity.py -m gpu -q --timeout 120 -x -k "tc_within" 2>&1 | tail -3
  timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_stage_parity.py -m gpu -q --timeout 100 -x -k "tc_within_tolerance_uniform_noise and 40-52" 2>&1 | grep -E "=========|passed|failed" | head -20
done
