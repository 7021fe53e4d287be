#!/usr/bin/env bash
# plan bench under two settings of an environment variable:  bash scripts/bench_ab.sh VAR v1 v2 [bench args]
var=$1; a=$2; b=$3; shift 3
for v in $a $b; do
  env $var=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_ab_${var}_$v.json 2> gpurun_out/bench_ab_${var}_$v.err || tail -3 gpurun_out/bench_ab_${var}_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_ab_${var}_$v.json"))
print("$var=$v", round(d["ms_per_step"], 4), "ms", round(d["value"], 3), {k.replace("pulpo_",""): round(x["ms_per_step"], 4) for k, x in d["kernels"].items()})
PY
done
