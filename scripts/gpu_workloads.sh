#!/usr/bin/env bash
# all bench workloads once (short runs) -> gpurun_out/wl_<tag>_*.json
tag=${1:-w}
out=gpurun_out
mkdir -p $out
for wl in oasis_4tot_4lat vecint_fullres; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $out/wl_${tag}_$wl.json 2> $out/wl_${tag}_$wl.err; echo "$wl rc=$?"; tail -2 $out/wl_${tag}_$wl.err
done
python bench.py --workload mc128 --samples ${SAMPLES:-16} --steps 2 --warmup 1 > $out/wl_${tag}_mc.json 2> $out/wl_${tag}_mc.err; echo "mc rc=$?"; tail -3 $out/wl_${tag}_mc.err
python - <<PY
import json
for wl in ("oasis_4tot_4lat","vecint_fullres","mc"):
    try:
        d=json.load(open("$out/wl_${tag}_%s.json"%wl))
        print(wl, round(d["value"],3), d["unit"], round(d["ms_per_step"],4),"ms", "e2e", round(d["e2e"]["value"],3), d.get("path_roofline",{}).get("frac"), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
        print("   ", {k.replace("pulpo_",""): round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e:
        print(wl, "failed", e)
PY
