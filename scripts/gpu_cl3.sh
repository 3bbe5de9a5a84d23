#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== timeline"
timeout 300 python scripts/timeline.py c2 2>&1 | tail -28 | tee gpurun_out/timeline.log
echo "== pingpong A/B"
for s in 0 1; do
  echo "pingpong=$s"; PMB_STREAM_MODE=3 PMB_CLUSTER_PINGPONG=$s timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-80
done | tee gpurun_out/ab.log
