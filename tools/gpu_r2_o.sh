#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/latency_breakdown.py > gpurun_out/r2o_latency.log 2>&1; cat gpurun_out/r2o_latency.log
CHI_B200_SPLIT_MAX=0 timeout 300 python tools/latency_breakdown.py > gpurun_out/r2o_latency_nosplit.log 2>&1; cat gpurun_out/r2o_latency_nosplit.log
