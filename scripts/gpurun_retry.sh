#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <timeout_s> <command...>   -- retries while the pod answers busy (nothing charged)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ]; then exit 0; fi
  echo "[retry $i] busy; sleeping 150 s"; sleep 150
done
