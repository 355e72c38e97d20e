#!/bin/bash
# round 2 final measurements on one GPU: profiles (default config and C2/1000) and the full bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
bash tools/gpu_profile.sh r2_final
bash tools/gpu_profile.sh r2_final_c2_1000 --config c2_1000
timeout 900 python bench.py > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_r2_final_reference.err
echo "reference exit $?"
