# gpurun -- 'VARIANTS="a b" bash scripts/bench_variants.sh'  : plan bench per library variant (kernel table only)
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export PULPO_B200_LIB=/root/repo/pulpo_b200/lib/libpulpo_b200_$v.so; else unset PULPO_B200_LIB; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_var_${v:-default}.json 2>gpurun_out/bench_var_${v:-default}.err || tail -3 gpurun_out/bench_var_${v:-default}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_var_${v:-default}.json"))
print("${v:-default}", round(d["ms_per_step"], 4), {k.replace("pulpo_",""): round(v["ms_per_step"], 4) for k, v in d["kernels"].items()})
PY
done
