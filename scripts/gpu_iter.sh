#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== timeline"
timeout 300 python scripts/timeline.py c2 2>&1 | tail -36 | tee gpurun_out/timeline.log
echo "== A/B"
for P in 2 4 8; do
  echo "P=$P"; PMB_PARTICLES_PER_CTA=$P timeout 300 python bench.py --steps 5 --warmup 3 --quick 2>&1 | tail -1
done | tee gpurun_out/ab.log
