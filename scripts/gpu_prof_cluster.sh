#!/bin/bash
# ncu full capture (with source) of the cluster-resident sweeps on the bench workload.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cluster_ -s 2 -c 2 -f -o gpurun_out/prof_cluster python scripts/profile_target.py c2 2 > gpurun_out/prof_cluster.log 2>&1
tail -2 gpurun_out/prof_cluster.log
ls -la gpurun_out/
