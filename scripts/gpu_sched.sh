#!/usr/bin/env bash
# compare plan scheduling options:  gpurun --timeout 900 -- 'bash scripts/gpu_sched.sh'
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_sched.log 2>&1
echo "pytest rc=$?"; tail -5 $out/pytest_sched.log
for cfg in "1 1 1" "1 1 0" "0 1 1" "0 1 0" "0 0 1" "1 0 1"; do
  set -- $cfg
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --fuse-combine $1 --pool-pyramid $2 --aux-early $3 > $out/bench_s$1$2$3.json 2> $out/bench_s$1$2$3.err
  python - <<PY
import json
d = json.load(open("$out/bench_s$1$2$3.json"))
print("combine=$1 pyramid=$2 aux_early=$3", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches_per_step"],
      {k.replace("pulpo_", ""): round(v["ms_per_step"], 4) for k, v in d["kernels"].items() if k in ("pulpo_avgpool2_pyramid_fwd", "pulpo_avgpool2_fwd", "pulpo_kl_n01_multi", "pulpo_combine_vecint_multi_fwd", "pulpo_combine_vecint_multi_bwd", "pulpo_vecint_multi_fwd", "pulpo_vecint_multi_bwd")})
PY
done
