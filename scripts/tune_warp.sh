for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export PULPO_B200_LIB=/root/repo/pulpo_b200/lib/libpulpo_b200_$v.so; else unset PULPO_B200_LIB; fi
  for m in ${MODES:-0 1}; do
    echo "== variant ${v:-default} mode $m"
    python scripts/prof_one.py warp 160 192 224 --mode $m
  done
done
