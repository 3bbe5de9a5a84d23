#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-220 | head -20 | tee gpurun_out/shapes.log
for c in c4 c5 c1; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --quick | tail -1; done | tee gpurun_out/other_configs.log
