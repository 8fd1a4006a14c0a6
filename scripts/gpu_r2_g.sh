#!/bin/bash
# 2-GPU: gradient identity of the sharded step (balanced on / off) + timing lines
mkdir -p gpurun_out
P=29511
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 10 --warmup 3 --check > gpurun_out/g_bench_2gpu_check_balanced.json 2> gpurun_out/g_2gpu_balanced.err
echo "rc=$?" >> gpurun_out/g_2gpu_balanced.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --steps 10 --warmup 3 --check --no-balance > gpurun_out/g_bench_2gpu_check_unbalanced.json 2> gpurun_out/g_2gpu_unbalanced.err
echo "rc=$?" >> gpurun_out/g_2gpu_unbalanced.err
ls -la gpurun_out
