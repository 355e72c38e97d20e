#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for ch in 1 4 8 16 32; do
CHI_B200_CHUNKS=$ch timeout 600 python bench.py --config c2 --no-extras --no-cpu-baseline --steps 5 > gpurun_out/r2s_ch$ch.json 2> gpurun_out/r2s_ch$ch.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2s_ch$ch.json')); r=d['roofline']
    print('chunks $ch value %.0f e2e %.0f (%.2f ms) dev ms %.2f'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['ms_per_step']))
except Exception as e: print('$ch ERR', e)
P
done
