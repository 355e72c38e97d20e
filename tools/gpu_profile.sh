# usage: bash tools/gpu_profile.sh <tag> [bench args...]   (on the GPU box; outputs under gpurun_out/)
tag=$1; shift
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-extras $*"
python bench.py $ARGS > gpurun_out/prof_${tag}_plain.json 2> gpurun_out/prof_${tag}_plain.err || { tail -3 gpurun_out/prof_${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${tag}.csv python bench.py $ARGS \
  > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on \
  -k regex:"solve_stream|solve_warp|observe_kernel" -s 2 -c 2 -f -o gpurun_out/prof_${tag} \
  python bench.py $ARGS > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out/prof_${tag}.ncu-rep gpurun_out/launches_${tag}.csv
