#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, smoke, short bench. Each leg has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
