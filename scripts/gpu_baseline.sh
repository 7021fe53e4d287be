#!/usr/bin/env bash
# One gpurun call: GPU parity tests, both bench arms, then the ncu launch list of the same
# bench command (profiling recipe: plain run first, ncu directly after with `&&`).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_baseline.sh r1c'
set -u
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log
python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "bench ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_$tag.log 2>&1
echo "ncu rc=$?"
python - <<EOF
import json
d = json.load(open("$out/bench_$tag.json"))
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "path_roofline", "gpu_launches_per_step")})
print({k: (round(v["ms_per_step"], 4), round(v["GBps"])) for k, v in d["kernels"].items()})
print(open("$out/bench_ref_$tag.json").read()[:600])
EOF
