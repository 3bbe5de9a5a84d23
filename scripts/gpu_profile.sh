#!/bin/bash
# ncu launch list + full capture of the sweep kernels (1 GPU) + default bench line.
mkdir -p gpurun_out
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python scripts/profile_target.py c2 3 > gpurun_out/launches.log 2>&1
tail -2 gpurun_out/launches.log
echo "== full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:cluster_(fwd|bwd)_kernel|rollout_" -s 2 -c 2 -f -o gpurun_out/prof_sweeps python scripts/profile_target.py c2 2 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
echo "== bench (default flags)"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.log
echo "== bench reference arm"
timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.log
