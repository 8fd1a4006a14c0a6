#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_shade_frontier.py > gpurun_out/c_shade_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/c_shade_ab.log
timeout 900 python -m pytest tests/test_gpu_dense.py::test_fused_csd_epilogue_matches_unfused tests/test_gpu_render.py tests/test_gpu_plugin.py -q -m gpu -s > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c_tests.log
timeout 1200 python -m pytest tests/test_gpu_config1.py -q -m gpu -s > gpurun_out/c_config1.log 2>&1
echo "config1 rc=$?" >> gpurun_out/c_config1.log
ls -la gpurun_out
