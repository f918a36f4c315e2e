#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-300; }
run t_gpu     python -m pytest tests -m gpu -q --timeout 300 -x
run b_tc      python bench.py --steps 50 --warmup 5 --no-cpu
python - <<'PY'
import json
l=[x for x in open('gpurun_out/b_tc.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('VALUE %.0f MPix/s  ms/step %.4f  e2e %.0f  tc_ms %.4f  frac %.3f  A_ms %.4f (%.2f hbm) C_ms %.4f (%.2f hbm) clocks %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['stages']['colour_bicubic_ms'],d['stages']['colour_bicubic_frac_hbm'],d['stages']['merge_ms'],d['stages']['merge_frac_hbm'],d['clocks']))
PY
