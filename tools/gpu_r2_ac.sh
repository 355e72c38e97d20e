#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/latency_one.py > gpurun_out/r2ac_lat_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"solve_warp" -s 3 -c 1 -f -o gpurun_out/prof_r2_final_single_vector python tools/latency_one.py > gpurun_out/r2ac_lat_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r2_final_single_vector.csv python tools/latency_one.py > gpurun_out/r2ac_lat_launches.log 2>&1
timeout 300 python tools/latency_breakdown.py > gpurun_out/r2ac_latency.log 2>&1; cat gpurun_out/r2ac_latency.log
ls -la gpurun_out/prof_r2_final_single_vector.ncu-rep
