#!/bin/bash
mkdir -p gpurun_out
timeout 120 prob_mbrl_b200/lib/umma_probe > gpurun_out/umma_probe.log 2>&1; echo rc=$?
wc -l gpurun_out/umma_probe.log
