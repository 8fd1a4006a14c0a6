#!/bin/bash
# host-GC A/B at the per-GPU work of the 8-GPU run (1 view on 1 GPU, 23 ms steps): 300 steps with / without the freeze
cd /root/repo
mkdir -p gpurun_out
for v in freeze nofreeze; do
  flag=""; [ $v = nofreeze ] && flag="--no-gc-freeze"
  timeout 400 python bench.py --views 1 --steps 300 --warmup 5 --no-cpu-baseline --no-gpu-baseline $flag > gpurun_out/i_1view_$v.json 2> gpurun_out/i_1view_$v.err
  echo "rc=$?" >> gpurun_out/i_1view_$v.err
done
python - <<'PY'
import json
for v in ("freeze","nofreeze"):
    d=json.loads(open(f"gpurun_out/i_1view_{v}.json").read().strip().splitlines()[-1])
    print(v, round(d["value"],2), d["step_time_spread"], "e2e", round(d["e2e"]["value"],2), d["e2e"]["step_time_spread"])
PY
