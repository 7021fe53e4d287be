#!/usr/bin/env bash
# quick GPU iteration: parity tests + bench (no CPU baseline, no ncu)
#   gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh tag'
set -u
tag=${1:-q}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -15 $out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches_per_step")}, d["path_roofline"]["frac"])
print({k.replace("pulpo_",""): (round(v["ms_per_step"], 4), round(v["GBps"])) for k, v in d["kernels"].items()})
PY
