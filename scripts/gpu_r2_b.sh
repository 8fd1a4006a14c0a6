#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_config1.py -q -m gpu -s > gpurun_out/b_config1.log 2>&1
echo "config1 rc=$?" >> gpurun_out/b_config1.log
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_system.py tests/test_gpu_render.py tests/test_gpu_plugin.py -q -m gpu -s > gpurun_out/b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b_tests.log
timeout 600 python scripts/exp_shade_frontier.py > gpurun_out/b_shade_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/b_shade_ab.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench rc=$?" >> gpurun_out/b_bench.err
ls -la gpurun_out
