#!/bin/bash
# Evidence run for the cluster-resident sweeps: timeline, sanitizers, other configs.
mkdir -p gpurun_out
export PMB_NO_PBAR=1
echo "== timeline"
timeout 300 python scripts/timeline.py c2 2>&1 | tail -28 | tee gpurun_out/timeline.log
echo "== memcheck (cluster-resident sweeps: parity tests on the small fixtures + mc_pilco)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "37x2 or mc_pilco_iterations" 2>&1 | tail -6 | tee gpurun_out/memcheck.log
echo "== racecheck (smoke = cluster-resident sweeps)"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/racecheck.log
echo "== synccheck (smoke)"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/synccheck.log
echo "== other configs (quick)"
for c in c1 c4 c5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --quick 2>&1 | tail -1
done | tee gpurun_out/other_configs.log
