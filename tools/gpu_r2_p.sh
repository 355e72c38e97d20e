#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1
tail -5 gpurun_out/r2p_pytest.log
python tools/latency_one.py > gpurun_out/r2p_lat_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"solve_warp" -s 3 -c 1 -f -o gpurun_out/prof_r2p_lat_s1 python tools/latency_one.py > gpurun_out/r2p_lat_ncu.log 2>&1
python tools/latency_one.py fwd > gpurun_out/r2p_lat_plain_fwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"solve_warp" -s 3 -c 1 -f -o gpurun_out/prof_r2p_lat_fwd python tools/latency_one.py fwd > gpurun_out/r2p_lat_ncu_fwd.log 2>&1
ls -la gpurun_out/prof_r2p*
