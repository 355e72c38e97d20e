python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ev.json 2>gpurun_out/ev.err || tail -3 gpurun_out/ev.err
python - <<PY
import json
d=json.load(open("gpurun_out/ev.json")); r=d["roofline"]
print(r["method"], "value %.0f e2e %.0f kernel_ms %.2f frac %.3f regs %d bpsm %d steps %.1f rej %.3f"%(d["value"], d["e2e"]["value"], r["kernel_ms"], r["frac"], r["registers"], r["blocks_per_sm"], r["steps_per_system"], r["rejected_fraction"]))
PY
