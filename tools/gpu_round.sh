#!/bin/bash
# Full GPU round (run under gpurun, 1 GPU): the whole gpu test-suite, smoke, bench (both arms), then the
# ncu launch list and one full capture of the fused kernel.  Logs and reports land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
run t_gpu     python -m pytest tests -m gpu -q --timeout 300
run smoke     python -c "import __graft_entry__ as g; g.smoke()"
run b_ref     python bench.py --impl reference --steps 2 --warmup 3
run b_tc      python bench.py --steps 50 --warmup 5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_srcnn_tc2 -s 3 -c 1 -f -o gpurun_out/prof_tc2 \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/prof_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_color_bicubic_tiled -s 3 -c 1 -f -o gpurun_out/prof_a \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/prof_bench.log 2>&1
echo done
