#!/usr/bin/env bash
# One gpurun call that regenerates a round's evidence set: GPU parity tests, smoke, both bench arms, the ncu
# launch list of the bench command and ONE `ncu --set full` pass over the launches of the first (eager) step.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_evidence.sh r2'
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log
python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "bench ref rc=$?"
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$cmd > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    $cmd > $out/ncu_$tag.log 2>&1
echo "ncu launches rc=$?"
# the first step runs eagerly: its launches are one of each kernel of the step, coarse levels first
ncu --set full --clock-control none --import-source on -k regex:'vecint_|ncc_tma_kernel|warp3d_|up2_|l2reg_fwd_bwd' -c 28 -f \
    -o $out/full_${tag} $cmd > $out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"; tail -2 $out/ncu_full_$tag.log
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "path_roofline", "gpu_launches_per_step")})
print({k.replace("pulpo_",""): (round(v["ms_per_step"], 4), round(v["GBps"])) for k, v in d["kernels"].items()})
print(open("$out/bench_ref_$tag.json").read()[:600])
PY
