#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --config c2_1000 --no-extras --no-cpu-baseline"
for cap in 4; do
CHI_B200_MAX_BLOCKS_PER_SM=$cap timeout 300 $B > gpurun_out/r2k_cap$cap.json 2> gpurun_out/r2k_cap$cap.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2k_cap$cap.json')); r=d['roofline']
    print('cap $cap value %.0f e2e %.0f frac %.3f kernel_ms %.3f regs %d bpsm %d steps %.1f'%(d['value'],d['e2e']['value'],r['frac'],r['kernel_ms'],r['registers'],r['blocks_per_sm'],r['steps_per_system']))
except Exception as e: print('$cap ERR', e)
P
done
bash tools/gpu_profile.sh r2_warp3 --config c2_1000
