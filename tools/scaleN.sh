# usage: bash tools/scaleN.sh <n>   (chains-sharded bench on n GPUs of one box)
n=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err || tail -5 gpurun_out/scale_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1]); print($n, "value %.0f e2e %.0f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
