#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== A/B"
for st in 2 3; do for P in 2 4 8; do
  echo "stages=$st P=$P"; PMB_STAGES=$st PMB_PARTICLES_PER_CTA=$P timeout 300 python bench.py --steps 5 --warmup 3 --quick 2>&1 | tail -1
done; done | tee gpurun_out/ab.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench.log
