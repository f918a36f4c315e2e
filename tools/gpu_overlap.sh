#!/bin/bash
# cross-call overlap check (run under gpurun): tests that exercise it, then same-box A/B of the bench step with the overlap off / on
# and over the merge kernel's grid size
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
[ -f tools/dbg_overlap.py ] && timeout 300 python tools/dbg_overlap.py 2>&1 | tail -12
timeout 900 python -m pytest tests/test_overlap.py tests/test_stage_parity.py tests/test_bands.py tests/test_host_pipeline.py -m gpu -q --timeout 600 2>&1 | tail -8
line() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stages']; print('$1','value %.0f ms/step %.4f ser %.4f | A %.4f B %.4f C %.4f between %s | e2e %.0f'%(d['value'],d['ms_per_step'],s['serialised_ms_per_step'],s['colour_bicubic_ms'],d['roofline']['kernel_ms'],s['merge_ms'],s['between_srcnn_launches_ms'],d['e2e']['value']))"; }
for r in 1 2; do
  SRCNN_OVERLAP=0 python bench.py --steps 50 --warmup 5 --no-cpu 2>gpurun_out/ov_err.log | line "overlap=0        "
  for m in ${CTAS:-2 4 5}; do
    SRCNN_MERGE_CTAS=$m python bench.py --steps 50 --warmup 5 --no-cpu 2>>gpurun_out/ov_err.log | line "overlap=1 ctas=$m "
  done
done
tail -5 gpurun_out/ov_err.log
