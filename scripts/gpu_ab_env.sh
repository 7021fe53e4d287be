#!/usr/bin/env bash
# A/B on the GPU over environment settings and bench flags: selected parity tests first, then one bench line per case.
#   gpurun --timeout 600 -- 'bash scripts/gpu_ab_env.sh TAG "<pytest -k expr>" "ENV=.. ENV2=.. -- <bench flags>" ...'
set -u
tag=$1; kexpr=$2; shift 2
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q -k "$kexpr" > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -4 $out/pytest_$tag.log
i=0
for c in "$@"; do
  envs=${c%%--*}; flags=${c#*--}
  env $envs python bench.py --steps 20 --warmup 5 --no-cpu-baseline $flags > $out/bench_${tag}_$i.json 2> $out/bench_${tag}_$i.err; echo "bench[$i] ($c) rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_${tag}_$i.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches_per_step")}, d["e2e"]["ms_per_step"], d["path_roofline"]["frac"])
    print({k.replace("pulpo_",""): (round(v["ms_per_step"], 4), round(v["GBps"])) for k, v in d["kernels"].items()})
except Exception as e:
    print("no bench line:", e); print(open("$out/bench_${tag}_$i.err").read()[-1500:])
PY
  i=$((i+1))
done
