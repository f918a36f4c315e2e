#!/bin/bash
# gpurun with retries while the pod is busy (exit 3 / "transient": nothing is charged).  usage: gpurun_retry.sh [gpurun args] -- cmd
for i in $(seq 1 40); do
  out=$(gpurun "$@" 2>&1); rc=$?
  echo "$out" | tail -n 60
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] busy, sleeping 90 s"; sleep 90; continue; fi
  exit $rc
done
exit 3
