#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1
tail -3 gpurun_out/r2y_pytest.log
timeout 300 python tools/latency_breakdown.py 2>&1 | head -8
