#!/bin/bash
# same-box A/B of the builds under build/alt/ (tools/build_alt.sh): parity tests on the default build first, then timings
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_stage_parity.py tests/test_bands.py tests/test_large_configs.py -m gpu -q --timeout 600 -x 2>&1 | tail -6
bash tools/ab_libs.sh 2>&1 | grep -v "^$" | cut -c1-220
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/b_tc.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/b_tc.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('VALUE %.0f MPix/s  ms/step %.4f  e2e %.0f (%.4f ms) tc_ms %.4f  frac %.3f  A_ms %.4f C_ms %.4f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['stages']['colour_bicubic_ms'],d['stages']['merge_ms']))
else: print(open('gpurun_out/b_tc.log').read()[-1500:])
PY
