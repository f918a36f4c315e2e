#!/bin/bash
# 1-GPU round: tests, smoke, bench lines of every configuration, parity report, JPEG stream numbers, ncu launch list and the
# --set full captures profiles/ is built from.  Logs land in gpurun_out/.   usage: gpu_round2.sh [noprof]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
js() { python - "$1" <<'PY'
import json,sys
l=[x for x in open('gpurun_out/%s.log'%sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}; s=d.get('stages') or {}
    print('%-16s N=%d value %.0f ms/step %.4f | e2e %s ms %s copy-only %s | K-B ms %s frac %s | A %s C %s | clocks %s'%(sys.argv[1],d['n_gpus'],d['value'],d['ms_per_step'],e.get('value'),e.get('ms_per_step'),e.get('copy_only_ms_per_step'),r.get('kernel_ms'),r.get('frac'),s.get('colour_bicubic_ms'),s.get('merge_ms'),(d.get('clocks') or {}).get('sm_mhz')))
else:
    print(open('gpurun_out/%s.log'%sys.argv[1]).read()[-1500:])
PY
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
TAILN=12 run t_gpu     python -m pytest tests -m gpu -q --timeout 900 -x
run smoke     python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 CUT=10 run b_cfg2 python bench.py --steps 50 --warmup 5; js b_cfg2
TAILN=1 CUT=10 run b_ref python bench.py --impl reference --steps 2 --warmup 3; js b_ref
TAILN=1 CUT=10 run b_cfg2_sustain python bench.py --steps 50 --warmup 5 --no-cpu --sustain 3; js b_cfg2_sustain
TAILN=1 CUT=10 run b_cfg3 python bench.py --config cfg3 --no-cpu; js b_cfg3
TAILN=1 CUT=10 run b_cfg4 python bench.py --config cfg4 --no-cpu; js b_cfg4
TAILN=1 CUT=10 run b_cfg5 python bench.py --config cfg5 --no-cpu; js b_cfg5
TAILN=3 CUT=600 run stream_1080 python tools/stream_bench.py 1920 1080 2.0 64
TAILN=3 CUT=600 run stream_4k_x4 python tools/stream_bench.py 3840 2160 4.0 16
TAILN=3 CUT=300 run parity python tools/parity_report.py
if [ "$1" != "noprof" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_srcnn_tc2 -s 3 -c 1 -f -o gpurun_out/prof_tc2 \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/prof_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_color_bicubic -s 3 -c 1 -f -o gpurun_out/prof_a \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/prof_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_merge_ycc2bgr -s 3 -c 1 -f -o gpurun_out/prof_c \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/prof_bench.log 2>&1
tail -2 gpurun_out/prof_bench.log
fi
echo done
