# final single-GPU measurements of a round: bash tools/final_run.sh <tag>
tag=$1
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || tail -3 gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err || tail -3 gpurun_out/bench_ref_${tag}.err
bash tools/gpu_profile.sh ${tag}
python bench.py --n-ids 10000 --chains 128 --no-cpu-baseline > gpurun_out/bench_10k_${tag}.json 2> gpurun_out/bench_10k_${tag}.err || tail -3 gpurun_out/bench_10k_${tag}.err
PYTHONPATH=. python profiles/r1_configs_script.py > gpurun_out/configs_${tag}.log 2>&1; tail -8 gpurun_out/configs_${tag}.log
python - <<PY
import json
for f in ("bench_${tag}", "bench_10k_${tag}"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "value %.0f e2e %.0f kernel_ms %.2f frac %.3f clocks %s cpu %s"%(d["value"], d["e2e"]["value"], r["kernel_ms"], r["frac"], d["clocks"], d.get("cpu_baseline",{}).get("value")))
d=json.loads(open("gpurun_out/bench_ref_${tag}.json").read().strip().splitlines()[-1]); print("ref", d["value"], d["cpu_baseline"]["cores"])
PY
