#!/bin/bash
# Round-2 evidence: launch list + full captures (c2 cluster sweeps, c3 / c5 streaming sweeps, c5 tensor-core sweeps).
mkdir -p gpurun_out
echo "== launch list (c2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python scripts/profile_target.py c2 3 > gpurun_out/launches.log 2>&1
tail -1 gpurun_out/launches.log
echo "== full capture c2"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:cluster_(fwd|bwd)_kernel" -s 2 -c 2 -f -o gpurun_out/prof_c2 python scripts/profile_target.py c2 2 > gpurun_out/prof_c2.log 2>&1
tail -1 gpurun_out/prof_c2.log
echo "== full capture c5 (wide cluster-resident sweeps)"
timeout 900 ncu --set full --clock-control none -k "regex:cw_(fwd|bwd)_kernel" -s 2 -c 2 -f -o gpurun_out/prof_c5 python scripts/profile_target.py c5 2 > gpurun_out/prof_c5.log 2>&1
tail -1 gpurun_out/prof_c5.log
echo "== full capture c4 (streaming sweeps)"
timeout 900 ncu --set full --clock-control none -k "regex:rollout_(fwd|bwd)" -s 2 -c 2 -f -o gpurun_out/prof_c4 python scripts/profile_target.py c4 2 > gpurun_out/prof_c4.log 2>&1
tail -1 gpurun_out/prof_c4.log
echo "== c3: ncu cannot replay the cooperative cluster launch of the moment-matching sweeps (empty report); launch list only"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv python scripts/profile_target.py c3 2 > gpurun_out/launches_c3.log 2>&1
tail -1 gpurun_out/launches_c3.log
echo "== full capture c5 (tensor-core sweeps, opt-in)"
PMB_STREAM_MODE=4 timeout 900 ncu --set full --clock-control none -k "regex:tc_(fwd|bwd)_kernel" -s 2 -c 2 -f -o gpurun_out/prof_c5tc python scripts/profile_target.py c5 2 > gpurun_out/prof_c5tc.log 2>&1
tail -1 gpurun_out/prof_c5tc.log
echo "== fit kernels launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_fit.csv python scripts/fit_target.py 4 > gpurun_out/launches_fit.log 2>&1
tail -1 gpurun_out/launches_fit.log
echo "== bench reference arm"
timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.log
echo "== bench (default flags)"
timeout 1200 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.log
