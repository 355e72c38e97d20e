#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2aa_pytest.log 2>&1
tail -3 gpurun_out/r2aa_pytest.log
for cfg in c2_1000 c2 c5 c4 c2_1000; do
timeout 600 python bench.py --config $cfg --no-extras --no-cpu-baseline > gpurun_out/r2aa_$cfg.json 2> gpurun_out/r2aa_$cfg.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2aa_$cfg.json')); r=d['roofline']
    print('$cfg value %.0f e2e %.0f (%.2f ms) frac %.3f kernel_ms %.3f'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],r['frac'],r['kernel_ms']))
except Exception as e: print('$cfg ERR', e)
P
done
