#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_shade_frontier.py > gpurun_out/e3_shade_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/e3_shade_ab.log
timeout 600 python -m pytest tests/test_gpu_render.py -q -m gpu -s > gpurun_out/e3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/e3_tests.log
