#!/bin/bash
mkdir -p gpurun_out
echo "== umma test"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "tensor_core" 2>&1 | tail -12 | tee gpurun_out/umma_test.log
echo "== A/B"
for u in 0 1; do
  echo "umma=$u"; PMB_WGRAD_UMMA=$u timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1
done | tee gpurun_out/umma_ab.log
echo "== ncu wgrad kernels"
for u in 0 1; do
  PMB_WGRAD_UMMA=$u timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wgrad -c 12 --csv --log-file gpurun_out/wgrad_umma$u.csv python scripts/profile_target.py c2 3 > /dev/null 2>&1
  awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/wgrad_umma$u.csv | cut -c1-40,150- | tail -6
done
