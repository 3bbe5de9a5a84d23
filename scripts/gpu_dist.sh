#!/bin/bash
# 2-GPU checks (tight timeouts: a hung collective must not burn the GPU budget):
# sharded == single-GPU parameters, then the weak-scaling bench line.
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | tail -2 | tee gpurun_out/dist_check.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_2gpu.log
