#!/bin/bash
# compute-sanitizer on the kernels added in round 2: wide cluster-resident sweeps (cw), moment matching inside the
# cluster-resident sweeps, tensor-core cluster sweeps, dynamics-model fit.  Small fixtures only (the tools slow kernels ~100x).
mkdir -p gpurun_out
export PMB_NO_PBAR=1
echo "== memcheck: cw / cluster mm / tc sweeps on the 7-particle fixture, fit trace"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fit.py -q -m gpu \
  -k "(37x2 and (cw or cluster or tc)) or golden_trace" 2>&1 | tail -6 | tee gpurun_out/r02_memcheck.log
echo "== racecheck: cw sweeps (7-particle fixture)"
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu \
  -k "test_rollout_and_gradient_match_reference_golden and 37x2 and cw" 2>&1 | tail -10 | tee gpurun_out/r02_racecheck_cw.log
echo "== racecheck: cluster-resident sweeps with moment matching (7-particle fixture)"
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu \
  -k "test_moment_matching_matches_reference_golden and 37x2 and cluster" 2>&1 | tail -10 | tee gpurun_out/r02_racecheck_mm.log
echo "== synccheck: cw sweeps + cluster mm"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu \
  -k "37x2 and (cw or (moment_matching_matches and cluster))" 2>&1 | tail -6 | tee gpurun_out/r02_synccheck.log
