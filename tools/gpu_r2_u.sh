#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CHI_B200_TIMELINE=1 python tools/e2e_timeline.py 4 12 2> gpurun_out/r2_timeline2.log; grep -A5 "chunks 4, call 2" gpurun_out/r2_timeline2.log; grep -A13 "chunks 12, call 2" gpurun_out/r2_timeline2.log
timeout 600 python tools/e2e_breakdown.py > gpurun_out/r2u_e2e.log 2>&1; cat gpurun_out/r2u_e2e.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1
tail -5 gpurun_out/r2u_pytest.log
