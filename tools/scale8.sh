python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err || tail -5 gpurun_out/scale_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n8.json").read().strip().splitlines()[-1]); print(8, "value %.0f e2e %.0f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
