#!/bin/bash
# same-box A/B of builds under build/alt/ that differ in the colour+bicubic kernel: bit-exact tests on the default build, then
# tools/ab_stages.py per build (two rounds)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_stage_parity.py tests/test_bands.py tests/test_random_geometries.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
for r in 1 2; do
  for lib in build/alt/lib_*.so; do
    echo "== $lib x2: $(SRCNN_B200_LIB=$PWD/$lib timeout 200 python tools/ab_stages.py 2>&1 | grep colour | cut -c1-120)"
  done
done
for lib in build/alt/lib_*.so; do
  echo "== $lib x4: $(SRCNN_B200_LIB=$PWD/$lib timeout 200 python tools/ab_stages.py 3840 2160 4 2>&1 | grep colour | cut -c1-120)"
done
