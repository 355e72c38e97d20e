#!/bin/bash
# new ROSL step (cached error weights, on-the-fly tail) + occupancy variants of the warp kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1
tail -3 gpurun_out/r2j_pytest.log
B="python bench.py --config c2_1000 --no-extras --no-cpu-baseline"
for tag in main w32x16 w32r120 w32r112 w32r104 w32x19 w32x21; do
lib=$PWD/chi_b200/libchi_b200_$tag.so; [ $tag = main ] && lib=$PWD/chi_b200/libchi_b200.so
CHI_B200_LIB=$lib timeout 300 $B > gpurun_out/r2j_$tag.json 2> gpurun_out/r2j_$tag.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2j_$tag.json')); r=d['roofline']
    print('$tag value %.0f e2e %.0f frac %.3f kernel_ms %.3f regs %d bpsm %d steps %.1f parity %s'%(d['value'],d['e2e']['value'],r['frac'],r['kernel_ms'],r['registers'],r['blocks_per_sm'],r['steps_per_system'],d.get('parity',{}).get('score_rel_err_vs_exact')))
except Exception as e: print('$tag ERR', e)
P
done
bash tools/gpu_profile.sh r2_warp2 --config c2_1000
timeout 300 python tools/latency_breakdown.py > gpurun_out/r2j_latency.log 2>&1; cat gpurun_out/r2j_latency.log
