#!/bin/bash
# 2-GPU round: the whole gpu suite (the multi-device tests run on two real devices), bench lines at N = 1 and 2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
nvidia-smi -L
TAILN=12 run t_gpu     python -m pytest tests -m gpu -q --timeout 900 -x
js() { python - "$1" <<'PY'
import json,sys
l=[x for x in open('gpurun_out/%s.log'%sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}; s=d.get('stages') or {}
    print('%-14s N=%d value %.0f  ms/step %.4f  e2e %s  K-B ms %s frac %s  A %s C %s wall %s'%(sys.argv[1],d['n_gpus'],d['value'],d['ms_per_step'],e.get('value'),r.get('kernel_ms'),r.get('frac'),s.get('colour_bicubic_ms'),s.get('merge_ms'),d.get('wall_ms_per_step')))
else:
    print(open('gpurun_out/%s.log'%sys.argv[1]).read()[-1500:])
PY
}
TAILN=1 CUT=10 run b_cfg3_n1 python bench.py --config cfg3 --no-cpu; js b_cfg3_n1
SRCNN_BATCH_LAUNCH=0 TAILN=1 CUT=10 run b_cfg3_n1_loop python bench.py --config cfg3 --no-cpu; js b_cfg3_n1_loop
TAILN=1 CUT=10 run b_cfg3_n2 python bench.py --config cfg3 --no-cpu --gpus 2; js b_cfg3_n2
TAILN=1 CUT=10 run b_cfg3_n2_mgpu python bench.py --config cfg3 --no-cpu --gpus 2 --mgpu; js b_cfg3_n2_mgpu
TAILN=1 CUT=10 run b_cfg4_n2 python bench.py --config cfg4 --no-cpu --gpus 2; js b_cfg4_n2
TAILN=1 CUT=10 run b_cfg4_n2_mgpu python bench.py --config cfg4 --no-cpu --gpus 2 --mgpu; js b_cfg4_n2_mgpu
TAILN=1 CUT=10 run b_cfg5_n2 python bench.py --config cfg5 --no-cpu --gpus 2; js b_cfg5_n2
TAILN=1 CUT=10 run b_cfg2_n2 python bench.py --no-cpu --gpus 2; js b_cfg2_n2
