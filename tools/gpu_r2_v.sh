#!/bin/bash
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CHI_B200_TIMELINE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2v_n${N}.json 2> gpurun_out/r2v_n${N}.err
python - <<P
import json
d=json.load(open('gpurun_out/r2v_n${N}.json'))
print('value %.0f e2e %.0f e2e ms %.3f dev ms %.3f'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['ms_per_step']))
P
grep "^chunk" gpurun_out/r2v_n${N}.err | tail -$((N*6)) 
