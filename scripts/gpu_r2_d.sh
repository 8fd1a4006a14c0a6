#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_config1.py tests/test_gpu_splitsum.py tests/test_gpu_render.py tests/test_gpu_dense.py tests/test_gpu_plugin.py tests/test_gpu_system.py -q -m gpu -s > gpurun_out/d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/d_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench rc=$?" >> gpurun_out/d_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --shading splitsum --no-cpu-baseline --no-gpu-baseline > gpurun_out/d_bench_splitsum.json 2> gpurun_out/d_bench_splitsum.err
echo "bench rc=$?" >> gpurun_out/d_bench_splitsum.err
timeout 600 ncu --set full --import-source on --section SourceCounters --clock-control none -k regex:shade_mc -c 1 -o gpurun_out/d_shade_full python scripts/prof_kernels.py shade > gpurun_out/d_ncu_shade.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"shade_splitsum|shade_bwd" -c 4 -o gpurun_out/d_splitsum_full python scripts/prof_kernels.py splitsum > gpurun_out/d_ncu_splitsum.log 2>&1
ls -la gpurun_out
