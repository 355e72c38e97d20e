#!/bin/bash
# warp-synchronous kernel: tests, A/B against the general kernel at three spreads, profile
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1
tail -3 gpurun_out/r2i_pytest.log
B="python bench.py --config c2_1000 --no-extras --no-cpu-baseline"
for ws in 1 0; do
for sp in 0.1 0.01 0.3; do
CHI_B200_WARP_SYNC=$ws timeout 300 $B --spread $sp > gpurun_out/r2i_ws${ws}_sp$sp.json 2> gpurun_out/r2i_ws${ws}_sp$sp.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2i_ws${ws}_sp$sp.json')); r=d['roofline']
    print('ws $ws spread $sp value %.0f e2e %.0f frac %.3f kernel_ms %.3f parity %s'%(d['value'],d['e2e']['value'],r['frac'],r['kernel_ms'],d.get('parity',{}).get('score_rel_err_vs_exact')))
except Exception as e: print('$ws $sp ERR', e)
P
done
done
bash tools/gpu_profile.sh r2_warp --config c2_1000
