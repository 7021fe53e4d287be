#!/usr/bin/env bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_full_prof.sh TAG'
# ncu --set full capture of every level-0 hot-path kernel from the bench command itself (plain run first)
set -u
tag=${1:-r1}
out=gpurun_out; mkdir -p $out
cmd="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$cmd > $out/plain_full_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'vecint_(fwd|bwd)_kernel' -c 2 -f -o $out/full_${tag}_vecint $cmd > $out/ncu_full_$tag.log 2>&1
echo "ncu vecint rc=$?"
# skip counts select the level-0 launches of the first (eager) step: levels run 3,2,1,0; levels 3 and 2 take the
# generic NCC kernel; up2 order: three combines, then the level-0 output resize forward and backward
ncu --set full --clock-control none --import-source on -k regex:'ncc_tma_kernel' -s 2 -c 2 -f -o $out/full_${tag}_ncc $cmd >> $out/ncu_full_$tag.log 2>&1
echo "ncu ncc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'warp3d_' -s 6 -c 2 -f -o $out/full_${tag}_warp $cmd >> $out/ncu_full_$tag.log 2>&1
echo "ncu warp rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'up2_' -s 3 -c 2 -f -o $out/full_${tag}_up2 $cmd >> $out/ncu_full_$tag.log 2>&1
echo "ncu up2 rc=$?"
tail -3 $out/ncu_full_$tag.log
