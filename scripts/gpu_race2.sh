#!/bin/bash
mkdir -p gpurun_out
export PMB_NO_PBAR=1
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "(200x2_n25_h40 and golden) or particles_per_cta or dcartpole_48x3_n24_h30-mmg" > gpurun_out/racecheck2_full.log 2>&1
grep -vE "^=========     and (Read|Write) access" gpurun_out/racecheck2_full.log | grep "=========" | cut -c1-260 | head -40
