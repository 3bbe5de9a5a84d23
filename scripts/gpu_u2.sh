#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_dimensional_actions" 2>&1 | tail -25 | tee gpurun_out/pytest_u2.log
