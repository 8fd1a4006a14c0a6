#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/z_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
echo "bench rc=$?" >> gpurun_out/z_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --shading splitsum --no-cpu-baseline --no-gpu-baseline > gpurun_out/z_bench_splitsum.json 2> gpurun_out/z_bench_splitsum.err
timeout 400 python bench.py --steps 20 --warmup 3 --views 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/z_bench_1view.json 2> gpurun_out/z_bench_1view.err
ls -la gpurun_out | tail -8
