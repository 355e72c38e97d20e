#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2ad_pytest.log 2>&1
tail -3 gpurun_out/r2ad_pytest.log
B="python bench.py --config c2_1000 --no-extras --no-cpu-baseline"
for sp in 0.1 0.3 0.01; do
for so in 1 0; do
CHI_B200_SORT=$so timeout 300 $B --spread $sp > gpurun_out/r2ad_so${so}_sp$sp.json 2> gpurun_out/r2ad_so${so}_sp$sp.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2ad_so${so}_sp$sp.json')); r=d['roofline']
    print('sort $so spread $sp value %.0f e2e %.0f frac %.3f kernel_ms %.3f step ms %.3f launches %s'%(d['value'],d['e2e']['value'],r['frac'],r['kernel_ms'],d['ms_per_step'],d['gpu_launches']))
except Exception as e: print('$so $sp ERR', e)
P
done
done
