#!/bin/bash
# host-GC A/B on the collective path: 2 GPUs, 1 view each (the per-GPU work of the 8-GPU run), 300 steps with / without the freeze
cd /root/repo
mkdir -p gpurun_out
for v in nofreeze freeze; do
  flag=""; [ $v = nofreeze ] && flag="--no-gc-freeze"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --views 2 --steps 300 --warmup 5 --no-cpu-baseline --no-gpu-baseline $flag > gpurun_out/j_2gpu_1view_$v.json 2> gpurun_out/j_2gpu_1view_$v.err
  echo "rc=$?" >> gpurun_out/j_2gpu_1view_$v.err
done
python - <<'PY'
import json
for v in ("freeze","nofreeze"):
    d=json.loads(open(f"gpurun_out/j_2gpu_1view_{v}.json").read().strip().splitlines()[-1])
    print(v, round(d["value"],2), d["step_time_spread"], "e2e", round(d["e2e"]["value"],2), d["e2e"]["step_time_spread"])
PY
