#!/bin/bash
mkdir -p gpurun_out
export PMB_NO_PBAR=1
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_full.log 2>&1
grep -E "hazard|Race reported|at .*cu|at .*cuh|Current Value|bytes" gpurun_out/racecheck_full.log | head -60
