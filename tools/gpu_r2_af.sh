#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for tag in main u8 main u8; do
lib=$PWD/chi_b200/libchi_b200_$tag.so; [ $tag = main ] && lib=$PWD/chi_b200/libchi_b200.so
CHI_B200_LIB=$lib timeout 300 python bench.py --config c4 --no-extras --no-cpu-baseline > gpurun_out/r2af_$tag.json 2> gpurun_out/r2af_$tag.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2af_$tag.json')); r=d['roofline']
    print('$tag c4 value %.0f e2e %.0f kernel_ms %.3f'%(d['value'],d['e2e']['value'],r['kernel_ms']))
except Exception as e: print('$tag ERR', e)
P
echo "$tag latency:"; CHI_B200_LIB=$lib timeout 300 python tools/latency_breakdown.py 2>&1 | grep -E "solve kernel|post\."
done
