#!/bin/bash
# multi-GPU check: the driver's launch line at N = $1
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n${N}_smi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_n${N}_bench.json 2> gpurun_out/r2_n${N}_bench.err
echo "exit $?" >> gpurun_out/r2_n${N}_bench.err
tail -5 gpurun_out/r2_n${N}_bench.err
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2_n${N}_bench.json'))
    print('value %.0f e2e %.0f ms %.3f scaling %s par %s shard_diff %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['scaling'],d['config']['parallelism'],d.get('sharded_vs_unsharded_score_rel_diff')))
    print('c2_1000', d['configs']['c2_1000']['value'], d['configs']['c2_1000']['parallelism'])
except Exception as e: print('ERR', e)
P
