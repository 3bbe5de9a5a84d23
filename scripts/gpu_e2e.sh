#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "mc_pilco or drop_in or loss_variants or default_flags" 2>&1 | tail -2 | tee gpurun_out/pytest_e2e.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_e2e.log
