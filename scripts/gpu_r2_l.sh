#!/bin/bash
# last check of the round: no cudaMalloc inside the step loop after reserve_step_scratch(); fused-step GPU tests still green
cd /root/repo
mkdir -p gpurun_out
timeout 100 python bench.py --views 1 --steps 300 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/l_1view.json 2> gpurun_out/l_1view.err
echo "rc=$?" >> gpurun_out/l_1view.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/l_1view.json").read().strip().splitlines()[-1])
print(round(d["value"],2), d["step_time_spread"], "e2e", round(d["e2e"]["value"],2), d["e2e"]["step_time_spread"])
PY
timeout 100 python -m pytest tests/test_gpu_system.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/l_tests.log
