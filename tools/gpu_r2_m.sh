#!/bin/bash
# validation of the round-2 kernels: GPU tests, full bench line, reference arm, profiles of the configs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1
tail -3 gpurun_out/r2m_pytest.log
timeout 900 python bench.py > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err
echo "bench exit $?"
bash tools/gpu_profile.sh r2m
bash tools/gpu_profile.sh r2m_c2_1000 --config c2_1000
bash tools/gpu_profile.sh r2m_c4 --config c4
bash tools/gpu_profile.sh r2m_c5 --config c5
timeout 300 python tools/latency_breakdown.py > gpurun_out/r2m_latency.log 2>&1; cat gpurun_out/r2m_latency.log
