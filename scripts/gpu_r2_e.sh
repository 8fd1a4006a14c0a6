#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_shade_frontier.py > gpurun_out/e_shade_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/e_shade_ab.log
timeout 600 python -m pytest tests/test_gpu_render.py -q -m gpu -s > gpurun_out/e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/e_tests.log
timeout 600 ncu --set full --import-source on --section SourceCounters --clock-control none -k regex:shade_mc -c 1 -o gpurun_out/e_shade_full python scripts/prof_kernels.py shade > gpurun_out/e_ncu_shade.log 2>&1
ls -la gpurun_out
