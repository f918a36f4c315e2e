#!/bin/bash
# compute-sanitizer memcheck over the parity, band, batch, host-pipeline and multi-worker tests (small shapes; kernels run ~20x slower)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SRCNN_WATCHDOG_MS=0 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_stage_parity.py tests/test_bands.py tests/test_host_pipeline.py tests/test_mgpu.py tests/test_jpeg_stream.py tests/test_overlap.py tests/test_color_bicubic_int.py -m gpu -q --timeout 1200 -x \
    > gpurun_out/sanitize.log 2>&1
echo "exit=$?" >> gpurun_out/sanitize.log
grep -E "ERROR SUMMARY|passed|failed|exit=|Invalid|out of bounds" gpurun_out/sanitize.log | head -20
