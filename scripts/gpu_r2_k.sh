#!/bin/bash
# collective path after the exchange warm-up: 2 GPUs, 1 view each, 300 steps; slowest_step diagnostics in the line
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --views 2 --steps 300 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/k_2gpu_1view.json 2> gpurun_out/k_2gpu_1view.err
echo "rc=$?" >> gpurun_out/k_2gpu_1view.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/k_2gpu_1view.json").read().strip().splitlines()[-1])
print(round(d["value"],2), d["step_time_spread"], "e2e", round(d["e2e"]["value"],2), d["e2e"]["step_time_spread"])
PY
tail -3 gpurun_out/k_2gpu_1view.err
