#!/bin/bash
# colour+bicubic kernel check (run under gpurun): bit-exact tests that exercise it, then same-box A/B of the integer-scale
# kernel against the generic tiled one (SRCNN_KA_INT=0), stage alone and inside the bench step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_stage_parity.py tests/test_bands.py tests/test_random_geometries.py -m gpu -q --timeout 600 -x 2>&1 | tail -6
for r in 1 2; do
  for v in 0 1; do
    echo "== SRCNN_KA_INT=$v x2"; SRCNN_KA_INT=$v timeout 200 python tools/ab_stages.py 2>&1 | grep colour
  done
done
for v in 0 1; do
  echo "== SRCNN_KA_INT=$v x4"; SRCNN_KA_INT=$v timeout 200 python tools/ab_stages.py 3840 2160 4 2>&1 | grep colour
  echo "== SRCNN_KA_INT=$v 720p x2"; SRCNN_KA_INT=$v timeout 200 python tools/ab_stages.py 1280 720 2 2>&1 | grep colour
done
for v in 0 1; do
SRCNN_KA_INT=$v python bench.py --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('KA_INT=$v','value %.0f A %.4f B %.4f C %.4f e2e %.0f'%(d['value'],d['stages']['colour_bicubic_ms'],d['roofline']['kernel_ms'],d['stages']['merge_ms'],d['e2e']['value']))"
done
