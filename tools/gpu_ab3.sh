#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 14 gpurun_out/$name.log | cut -c1-400; }
run t_par     python -m pytest tests/test_stage_parity.py tests/test_bands.py -m gpu -q --timeout 120 -x -k "tc or band"
run ab        python tools/ab_tc2.py
timeout 120 python tools/tc2_timeline.py > gpurun_out/timeline_rot.log 2>&1; tail -48 gpurun_out/timeline_rot.log
