#!/bin/bash
# 8-GPU: north-star strong-scaling line (8 views, 1/GPU) with --check, config 4 weak-scaling line (64 views, 8/GPU)
mkdir -p gpurun_out
P=29611
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 3 --check > gpurun_out/h_bench_8gpu.json 2> gpurun_out/h_8gpu.err
echo "rc=$?" >> gpurun_out/h_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 8 --steps 10 --warmup 3 --views 64 > gpurun_out/h_bench_8gpu_config4_64views.json 2> gpurun_out/h_8gpu_c4.err
echo "rc=$?" >> gpurun_out/h_8gpu_c4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/h_bench_4gpu.json 2> gpurun_out/h_4gpu.err
echo "rc=$?" >> gpurun_out/h_4gpu.err
ls -la gpurun_out
