#!/bin/bash
# Runs on the GPU box (via gpurun): each stage in its own process and under its own timeout, so one
# trapped kernel cannot poison the rest.  Logs go to gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 15 gpurun_out/$name.log; }
run t_prims   python -m pytest tests/test_tc_primitives.py -m gpu -q --timeout 120
run t_nontc   python -m pytest tests/test_stage_parity.py tests/test_bands.py -m gpu -q --timeout 300 -k "not tc"
run t_tc      python -m pytest tests/test_stage_parity.py tests/test_bands.py -m gpu -q --timeout 300 -k "tc"
run smoke     python -c "import __graft_entry__ as g; g.smoke()"
run b_fp32    python bench.py --steps 10 --warmup 3 --variant fp32 --no-cpu
run b_tc      python bench.py --steps 50 --warmup 5 --no-cpu
