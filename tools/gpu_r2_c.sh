#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --config c2_1000 --no-extras --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2c_base.json 2> gpurun_out/r2c_base.err
CHI_B200_REFILL_WAIT=0 timeout 300 $B > gpurun_out/r2c_wait0.json 2> gpurun_out/r2c_wait0.err
CHI_B200_REFILL_WAIT=8 timeout 300 $B > gpurun_out/r2c_wait8.json 2> gpurun_out/r2c_wait8.err
CHI_B200_ORDER=0 timeout 300 $B > gpurun_out/r2c_order0.json 2> gpurun_out/r2c_order0.err
timeout 300 $B --spread 0.01 > gpurun_out/r2c_spread001.json 2> gpurun_out/r2c_spread001.err
timeout 300 $B --spread 0.3 > gpurun_out/r2c_spread03.json 2> gpurun_out/r2c_spread03.err
for f in base wait0 wait8 order0 spread001 spread03; do python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2c_$f.json')); r=d['roofline']
    print('$f value %.0f e2e %.0f frac %.3f kernel_ms %.3f steps %.1f'%(d['value'],d['e2e']['value'],r['frac'],r['kernel_ms'],r['steps_per_system']))
except Exception as e: print('$f ERR', e)
P
done
