#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 3 --check > gpurun_out/h2_bench_8gpu.json 2> gpurun_out/h2_8gpu.err
echo "rc=$?" >> gpurun_out/h2_8gpu.err
