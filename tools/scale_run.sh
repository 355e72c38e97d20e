# 1 -> 8 GPU scaling of bench.py on one box (chains sharded) + individuals-sharded 10k model on 8
set -x
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -c 300 gpurun_out/scale_n$n.json | head -c 300; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --shard individuals --n-ids 1250 > gpurun_out/scale_n8_individuals.json 2> gpurun_out/scale_n8_individuals.err
python - <<PY
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/scale_n%d.json"%n).read().strip().splitlines()[-1]); print(n, "value %.0f e2e %.0f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]))
    except Exception as e: print(n, "ERR", e)
try:
    d=json.loads(open("gpurun_out/scale_n8_individuals.json").read().strip().splitlines()[-1]); print("ind", d["value"], d["e2e"]["value"], d["config"]["parallelism"], d["config"]["n_ids"])
except Exception as e: print("ind ERR", e)
PY
