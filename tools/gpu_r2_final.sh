#!/bin/bash
# round 2 final measurements on one GPU: tests, the full bench line, the reference arm, profiles of the configs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1
tail -3 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_r2_final_reference.err
echo "reference exit $?"
bash tools/gpu_profile.sh r2_final
bash tools/gpu_profile.sh r2_final_c2_1000 --config c2_1000
bash tools/gpu_profile.sh r2_final_c4 --config c4
bash tools/gpu_profile.sh r2_final_c5 --config c5
