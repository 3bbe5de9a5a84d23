#!/bin/bash
mkdir -p gpurun_out
echo "== parity (cluster subset)"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "particles_per_cluster or sweep_variants or golden or cotangents" 2>&1 | tail -3 | tee gpurun_out/cluster_small.log
echo "== A/B flags"
for f in 1 3; do
  echo "flags=$f"; PMB_STREAM_MODE=3 PMB_CLUSTER_PINGPONG=$f timeout 200 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-90
done | tee gpurun_out/ab.log
