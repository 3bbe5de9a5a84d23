#!/bin/bash
# First contact of the cluster-resident sweeps with a B200: small parity test, sanitizer, full tests, timeline, bench.
mkdir -p gpurun_out
echo "== small parity (cluster)"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "particles_per_cluster or sweep_variants" 2>&1 | tail -15 | tee gpurun_out/cluster_small.log
echo "== memcheck (cluster, tiny)"
PMB_STREAM_MODE=3 timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/cluster_memcheck.log
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== timeline"
timeout 300 python scripts/timeline.py c2 2>&1 | tail -30 | tee gpurun_out/timeline.log
echo "== A/B"
for m in 2 3; do
  echo "mode=$m"; PMB_STREAM_MODE=$m timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1
done | tee gpurun_out/ab.log
for c in "7 8" "8 8" "4 4" "3 4"; do set -- $c
  echo "PG=$1 C=$2"; PMB_STREAM_MODE=3 PMB_CLUSTER_PG=$1 PMB_CLUSTER_C=$2 timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1
done | tee -a gpurun_out/ab.log
