#!/bin/bash
mkdir -p gpurun_out
echo "== timeline with one particle per cluster (second tile idle: no inter-group contention)"
PMB_STREAM_MODE=3 PMB_CLUSTER_PG=1 timeout 300 python scripts/timeline.py c2 2>&1 | tail -28 | cut -c1-110 | tee gpurun_out/timeline_pg1.log
echo "== PG=2 (1+1)"
PMB_STREAM_MODE=3 PMB_CLUSTER_PG=2 timeout 300 python scripts/timeline.py c2 2>&1 | tail -28 | head -16 | cut -c1-110 | tee gpurun_out/timeline_pg2.log
