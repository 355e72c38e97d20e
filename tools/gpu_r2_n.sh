#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1
tail -30 gpurun_out/r2n_pytest.log
timeout 300 python tools/latency_breakdown.py > gpurun_out/r2n_latency.log 2>&1; cat gpurun_out/r2n_latency.log
CHI_B200_SPLIT_MAX=0 timeout 300 python tools/latency_breakdown.py > gpurun_out/r2n_latency_nosplit.log 2>&1; cat gpurun_out/r2n_latency_nosplit.log
