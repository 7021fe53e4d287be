# gpurun -- 'bash scripts/tune_vecint.sh'   (variants built with python -m pulpo_b200.build --tag=...)
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export PULPO_B200_LIB=/root/repo/pulpo_b200/lib/libpulpo_b200_$v.so; else unset PULPO_B200_LIB; fi
  for m in ${MODES:-0 2 0x102}; do
    echo "== variant ${v:-default} mode $m"
    python scripts/prof_one.py vecint 80 96 112 --mode $m
  done
done
