for v in "" e1 e2; do
  if [ -n "$v" ]; then export PULPO_B200_LIB=/root/repo/pulpo_b200/lib/libpulpo_b200_$v.so; else unset PULPO_B200_LIB; fi
  echo "== variant ${v:-default}"
  python scripts/prof_one.py vecint 80 96 112
done
