#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_prof.sh <tag> <kernel-regex> <prof_one args...>'
# plain timing of the op at all hot-path shapes, then one ncu --set full capture of the named kernel
set -u
tag=$1; regex=$2; shift 2
out=gpurun_out; mkdir -p $out
python scripts/prof_one.py "$@" > $out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 2 -f -o $out/prof_$tag \
    python scripts/prof_one.py "$@" --reps 2 > $out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; cat $out/plain_$tag.log; tail -3 $out/ncu_$tag.log
