#!/bin/bash
# GPU call 1 of the session: parity tests of everything touched + knob A/B + one bench line + a timeline
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*" ; timeout 700 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 14 gpurun_out/$name.log | cut -c1-600; }
run t_par     python -m pytest tests/test_tc_primitives.py tests/test_stage_parity.py tests/test_bands.py -m gpu -q --timeout 300 -x
run ab        python tools/ab_tc2.py
run b_tc      python bench.py --steps 50 --warmup 5 --no-cpu
SRCNN_TC2_E1_WIDE=0 timeout 120 python tools/tc2_timeline.py > gpurun_out/timeline_e1narrow.log 2>&1; tail -3 gpurun_out/timeline_e1narrow.log
SRCNN_TC2_E1_WIDE=1 timeout 120 python tools/tc2_timeline.py > gpurun_out/timeline_e1wide.log 2>&1; tail -3 gpurun_out/timeline_e1wide.log
run t_rest    python -m pytest tests -m gpu -q --timeout 300 --ignore tests/test_stage_parity.py --ignore tests/test_tc_primitives.py --ignore tests/test_bands.py
echo done
