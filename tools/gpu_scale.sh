#!/bin/bash
# Multi-GPU round on ONE box with NG GPUs (gpurun --gpus NG): bench lines of every configuration at N = NG, one rank per GPU
# (torchrun) and through the in-library driver (--mgpu, one process); at NG = 8 also the host-buffer knobs.
NG=${1:-8}
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
nvidia-smi -L | head -8; nproc; free -g | head -2
run b_cfg2_n$NG python bench.py --no-cpu --gpus $NG
run b_cfg3_n$NG python bench.py --config cfg3 --no-cpu --gpus $NG
run b_cfg4_n$NG python bench.py --config cfg4 --no-cpu --gpus $NG
run b_cfg5_n$NG python bench.py --config cfg5 --no-cpu --gpus $NG
run b_cfg3_n${NG}_mgpu python bench.py --config cfg3 --no-cpu --gpus $NG --mgpu
run b_cfg4_n${NG}_mgpu python bench.py --config cfg4 --no-cpu --gpus $NG --mgpu
run b_cfg5_n${NG}_mgpu python bench.py --config cfg5 --no-cpu --gpus $NG --mgpu
if [ "$NG" = "8" ]; then
  SRCNN_HOST_BANDS=2 run b_cfg2_n8_bands2 python bench.py --no-cpu --gpus 8 --steps 20
  SRCNN_HOST_BANDS=4 run b_cfg2_n8_bands4 python bench.py --no-cpu --gpus 8 --steps 20
  SRCNN_GRAPHS=0 run b_cfg2_n8_nograph python bench.py --no-cpu --gpus 8 --steps 20
fi
