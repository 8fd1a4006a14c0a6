#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?" >> gpurun_out/f_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --views 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/f_bench_config2.json 2> gpurun_out/f_bench_config2.err
timeout 900 python bench.py --steps 10 --warmup 3 --res 1024 --dtype bf16 --no-cpu-baseline --no-gpu-baseline > gpurun_out/f_bench_config3.json 2> gpurun_out/f_bench_config3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
ls -la gpurun_out
