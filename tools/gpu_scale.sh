#!/bin/bash
# Multi-GPU round on ONE box with NG GPUs (gpurun --gpus NG): bench lines of every configuration at N = NG, one rank per GPU
# (torchrun) and through the in-library driver (--mgpu, one process).  usage: gpu_scale.sh NG [cfgs...]
NG=${1:-8}; shift
CFGS=${@:-cfg2 cfg3 cfg4 cfg5}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; python - "$name" <<'PY'
import json,sys
l=[x for x in open('gpurun_out/%s.log'%sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); r=d.get('roofline') or {}; e=d.get('e2e') or {}
    print('%-18s N=%d value %.0f ms/step %.4f | e2e %s ms %s copy-only ms %s ceilGB/s %s | K-B frac %s | wall %s'%(sys.argv[1],d['n_gpus'],d['value'],d['ms_per_step'],e.get('value'),e.get('ms_per_step'),e.get('copy_only_ms_per_step'),e.get('pcie_ceiling_GBs'),r.get('frac'),d.get('wall_ms_per_step')))
else:
    print(sys.argv[1], open('gpurun_out/%s.log'%sys.argv[1]).read()[-1200:])
PY
}
nvidia-smi -L | wc -l; nproc
for c in $CFGS; do
  run b_${c}_n$NG python bench.py --config $c --no-cpu --gpus $NG
  if [ "$c" != "cfg2" ] && [ -z "$NOMGPU" ]; then run b_${c}_n${NG}_mgpu python bench.py --config $c --no-cpu --gpus $NG --mgpu; fi
done
