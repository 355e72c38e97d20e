#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/e2e_breakdown.py > gpurun_out/r2t_e2e.log 2>&1; cat gpurun_out/r2t_e2e.log
