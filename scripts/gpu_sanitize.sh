#!/bin/bash
mkdir -p gpurun_out
export PMB_NO_PBAR=1
echo "== memcheck (smoke + mm + mc_pilco)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "37x2 or mc_pilco or rank_deficient" 2>&1 | tail -8 | tee gpurun_out/memcheck.log
echo "== racecheck (smoke)"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/racecheck.log
echo "== synccheck (smoke)"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/synccheck.log
