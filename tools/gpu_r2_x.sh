#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1
tail -3 gpurun_out/r2x_pytest.log
for cfg in c5; do
timeout 600 python bench.py --config $cfg --no-extras --no-cpu-baseline > gpurun_out/r2x_$cfg.json 2> gpurun_out/r2x_$cfg.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2x_$cfg.json')); r=d['roofline']
    print('$cfg value %.0f e2e %.0f (%.2f ms) frac %.3f kernel_ms %.3f %s'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],r['frac'],r['kernel_ms'],r['kernel'][:40]))
except Exception as e: print('$cfg ERR', e)
P
done
