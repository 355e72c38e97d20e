#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/latency_breakdown.py > gpurun_out/r2q_latency.log 2>&1; cat gpurun_out/r2q_latency.log
timeout 600 python -m pytest tests -m gpu -q -k "small_launches or bench_tolerance" 2>&1 | tail -3
