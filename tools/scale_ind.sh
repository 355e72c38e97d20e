python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --shard individuals --n-ids 1250 > gpurun_out/scale_n2_individuals.json 2> gpurun_out/scale_n2_individuals.err || tail -5 gpurun_out/scale_n2_individuals.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n2_individuals.json").read().strip().splitlines()[-1]); print("ind", d["value"], d["e2e"], d["config"]["parallelism"])
PY
