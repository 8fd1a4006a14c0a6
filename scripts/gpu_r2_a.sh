#!/bin/bash
# round-2 GPU call A: new hp / J1 / full-size parity tests, config-1 run, shader source-level profile, quick bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_hp.py tests/test_gpu_dense.py -q -m gpu -s -x --deselect tests/test_gpu_hp.py::test_hp_full_size_vae_512_and_controlnet_64 > gpurun_out/a_tests1.log 2>&1
echo "tests1 rc=$?" >> gpurun_out/a_tests1.log
timeout 900 python -m pytest tests/test_gpu_hp.py::test_hp_full_size_vae_512_and_controlnet_64 -q -m gpu -s > gpurun_out/a_tests2.log 2>&1
echo "tests2 rc=$?" >> gpurun_out/a_tests2.log
timeout 1200 python -m pytest tests/test_gpu_config1.py -q -m gpu -s > gpurun_out/a_config1.log 2>&1
echo "config1 rc=$?" >> gpurun_out/a_config1.log
timeout 600 python -m pytest tests/test_gpu_system.py tests/test_gpu_render.py -q -m gpu -s > gpurun_out/a_tests3.log 2>&1
echo "tests3 rc=$?" >> gpurun_out/a_tests3.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 900 ncu --set full --import-source on --section SourceCounters --section WarpStateStats --clock-control none -k regex:shade_mc -c 1 -o gpurun_out/a_shade_full python scripts/prof_kernels.py shade > gpurun_out/a_ncu_shade.log 2>&1
echo "ncu rc=$?" >> gpurun_out/a_ncu_shade.log
ls -la gpurun_out
