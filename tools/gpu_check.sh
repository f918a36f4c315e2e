#!/bin/bash
# GPU check (run under gpurun, 1 GPU): the whole gpu test-suite, smoke, one bench line.  Logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-600; }
run t_gpu     python -m pytest tests -m gpu -q --timeout 600 -x
run smoke     python -c "import __graft_entry__ as g; g.smoke()"
TAILN=3 run b_tc      python bench.py --steps 50 --warmup 5 --no-cpu
python - <<'PY'
import json
l=[x for x in open('gpurun_out/b_tc.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('VALUE %.0f MPix/s  ms/step %.4f  e2e %.0f  tc_ms %.4f  frac %.3f  A_ms %.4f C_ms %.4f clocks %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['stages']['colour_bicubic_ms'],d['stages']['merge_ms'],d['clocks']))
PY
