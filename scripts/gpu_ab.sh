#!/usr/bin/env bash
# quick A/B on the GPU: selected parity tests, then bench lines for a list of flag sets
#   gpurun --timeout 900 -- 'bash scripts/gpu_ab.sh TAG "<pytest -k expr>" "<flags A>" "<flags B>" ...'
set -u
tag=$1; kexpr=$2; shift 2
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q -k "$kexpr" > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -12 $out/pytest_$tag.log
i=0
for flags in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline $flags > $out/bench_${tag}_$i.json 2> $out/bench_${tag}_$i.err; echo "bench[$i] ($flags) rc=$?"
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
