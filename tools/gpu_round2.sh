#!/bin/bash
# 1-GPU round: tests, smoke, bench lines of every configuration, parity report.  Logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=12 run t_gpu     python -m pytest tests -m gpu -q --timeout 900 -x
run smoke     python -c "import __graft_entry__ as g; g.smoke()"
CUT=3000 TAILN=2 run b_cfg2 python bench.py --steps 50 --warmup 5 --no-cpu
CUT=3000 TAILN=2 run b_cfg2_sustain python bench.py --steps 50 --warmup 5 --no-cpu --sustain 3
CUT=3000 TAILN=2 run b_cfg3 python bench.py --config cfg3 --no-cpu
CUT=3000 TAILN=2 run b_cfg4 python bench.py --config cfg4 --no-cpu
CUT=3000 TAILN=2 run b_cfg5 python bench.py --config cfg5 --no-cpu
CUT=600 TAILN=8 run parity python tools/parity_report.py
