#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 2 -c 2 -f -o gpurun_out/prof_sweeps python scripts/profile_target.py c2 2 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
