#!/bin/bash
# ncu launch list + full capture of the sweep kernels (1 GPU), plus A/B of tunables.
mkdir -p gpurun_out
echo "== A/B"
for mode in 1 2; do for P in 2 4 8; do
  echo "stream_mode=$mode P=$P"; PMB_STREAM_MODE=$mode PMB_PARTICLES_PER_CTA=$P timeout 300 python bench.py --steps 5 --warmup 3 --quick 2>&1 | tail -1
done; done | tee gpurun_out/ab.log
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python scripts/profile_target.py c2 3 > gpurun_out/launches.log 2>&1
tail -3 gpurun_out/launches.log
echo "== full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 2 -c 2 -f -o gpurun_out/prof_sweeps python scripts/profile_target.py c2 2 > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
ls -la gpurun_out
