# usage: bash tools/gpu_sweep.sh "<bench args 1>" "<bench args 2>" ...
i=0
for args in "$@"; do
  i=$((i+1))
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline $args > gpurun_out/sw_$i.json 2>gpurun_out/sw_$i.err || tail -3 gpurun_out/sw_$i.err
  python - "$args" gpurun_out/sw_$i.json <<PY
import json,sys
d=json.load(open(sys.argv[2])); r=d["roofline"]
print(sys.argv[1], "| value %.0f kernel_ms %.2f frac %.3f steps %.1f rej %.3f"%(d["value"], r["kernel_ms"], r["frac"], r["steps_per_system"], r["rejected_fraction"]))
PY
done
