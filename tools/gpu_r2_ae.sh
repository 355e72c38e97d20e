#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --config c2_1000 --no-extras --no-cpu-baseline"
for tag in main rhsd u8 main rhsd; do
lib=$PWD/chi_b200/libchi_b200_$tag.so; [ $tag = main ] && lib=$PWD/chi_b200/libchi_b200.so
CHI_B200_LIB=$lib timeout 300 $B > gpurun_out/r2ae_$tag.json 2> gpurun_out/r2ae_$tag.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2ae_$tag.json')); r=d['roofline']
    print('$tag value %.0f e2e %.0f frac %.3f kernel_ms %.3f steps %.2f'%(d['value'],d['e2e']['value'],r['frac'],r['kernel_ms'],r['steps_per_system']))
except Exception as e: print('$tag ERR', e)
P
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
