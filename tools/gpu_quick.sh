#!/bin/bash
# quick GPU loop: primitives + TC parity + bench (no ncu)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log; }
run t_tc      python -m pytest tests/test_tc_primitives.py tests/test_stage_parity.py tests/test_bands.py -m gpu -q --timeout 300 -x
run b_tc      python bench.py --steps 50 --warmup 5 --no-cpu
python - <<'PY'
import json
l=[x for x in open('gpurun_out/b_tc.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('VALUE %.0f MPix/s  ms/step %.4f  e2e %.0f  tc_ms %.4f  frac %.3f  A_ms %.4f C_ms %.4f clocks %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['stages']['colour_bicubic_ms'],d['stages']['merge_ms'],d['clocks']))
PY
timeout 120 python tools/tc2_timeline.py 2>&1 | tail -60
