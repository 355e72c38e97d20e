#!/bin/bash
# round 2: ROSL(11) on the GPU -- parity suite, variant comparison, the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
for v in 0 1 2; do
  timeout 300 python bench.py --config c2_1000 --no-extras --no-cpu-baseline --variant $v > gpurun_out/r2b_c2_1000_v$v.json 2> gpurun_out/r2b_c2_1000_v$v.err
done
timeout 300 python bench.py --config c2_1000 --no-extras --no-cpu-baseline --rtol 1e-6 --atol 1e-8 > gpurun_out/r2b_c2_1000_v0_tol6.json 2> gpurun_out/r2b_c2_1000_v0_tol6.err
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench exit $?" >> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_pytest.log
