#!/bin/bash
# round 2, first GPU check: parity suite under the new defaults, the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit $?" >> gpurun_out/r2a_bench.err
timeout 300 python bench.py --config c2_1000 --no-extras --no-cpu-baseline --spread 0.01 > gpurun_out/r2a_c2_1000_fused.json 2> gpurun_out/r2a_c2_1000_fused.err
CHI_B200_SNAP_MB=100000 timeout 300 python bench.py --config c2_1000 --no-extras --no-cpu-baseline --spread 0.01 > gpurun_out/r2a_c2_1000_deferred.json 2> gpurun_out/r2a_c2_1000_deferred.err
tail -3 gpurun_out/r2a_pytest.log
