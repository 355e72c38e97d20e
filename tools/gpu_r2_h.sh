#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
B="python bench.py --config c2_1000 --no-extras --no-cpu-baseline"
for sp in 0.1 0.01 0.3; do
timeout 300 $B --spread $sp > gpurun_out/r2h_sp$sp.json 2> gpurun_out/r2h_sp$sp.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2h_sp$sp.json')); r=d['roofline']
    print('spread $sp value %.0f e2e %.0f frac %.3f kernel_ms %.3f'%(d['value'],d['e2e']['value'],r['frac'],r['kernel_ms']))
except Exception as e: print('$sp ERR', e)
P
done
bash tools/gpu_profile.sh r2_rosl_c --config c2_1000
