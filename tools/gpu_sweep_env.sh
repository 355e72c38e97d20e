# usage: bash tools/gpu_sweep_env.sh "<ENV=.. ENV=..|bench args>" ...   (each arg: env assignments, then '|' and bench args)
i=0
for spec in "$@"; do
  i=$((i+1))
  envs="${spec%%|*}"; args="${spec#*|}"
  env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline $args > gpurun_out/sw_$i.json 2>gpurun_out/sw_$i.err || tail -3 gpurun_out/sw_$i.err
  python - "$spec" gpurun_out/sw_$i.json <<PY
import json,sys
d=json.load(open(sys.argv[2])); r=d["roofline"]
print(sys.argv[1], "| value %.0f e2e %.0f kernel_ms %.2f frac %.3f steps %.1f rej %.3f"%(d["value"], d["e2e"]["value"], r["kernel_ms"], r["frac"], r["steps_per_system"], r["rejected_fraction"]))
PY
done
