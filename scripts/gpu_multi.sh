#!/usr/bin/env bash
# multi-GPU measurements on ONE box:  gpurun --gpus 8 -- 'bash scripts/gpu_multi.sh tag'
tag=${1:-m}
out=gpurun_out
mkdir -p $out
ng=$(nvidia-smi -L | wc -l)
run() {  # n, name, args...
  n=$1; name=$2; shift 2
  if [ "$n" -gt "$ng" ]; then return; fi
  if [ "$n" -eq 1 ]; then
    python bench.py --gpus 1 "$@" > $out/${tag}_${name}_n$n.json 2> $out/${tag}_${name}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > $out/${tag}_${name}_n$n.json 2> $out/${tag}_${name}_n$n.err
  fi
  echo "$name n=$n rc=$?"; tail -2 $out/${tag}_${name}_n$n.err | cut -c1-300
}
python -m pytest tests/test_gpu_multirank.py -q 2>&1 | tail -3
for n in 1 2 4 8; do run $n mc128 --workload mc128 --steps 3 --warmup 1; done
for n in 1 2 4 8; do run $n pairs --steps 20 --warmup 5 --no-cpu-baseline; done
for n in 2 4 8; do run $n cfg5 --batch $((32 / n)) --grad-allreduce 1 --steps 5 --warmup 2 --no-cpu-baseline; done
python - <<PY
import json, glob
for name in ("mc128", "pairs", "cfg5"):
    base = None
    for n in (1, 2, 4, 8):
        try:
            d = json.load(open("$out/${tag}_%s_n%d.json" % (name, n)))
        except Exception as e:
            continue
        v = d["value"]
        if base is None:
            base = (n, v)
        print(name, "n=%d" % n, "value %.3f %s" % (v, d["unit"]), "ms/step %.4f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["value"],
              "eff vs n=%d: %.3f" % (base[0], v / base[1] / (n / base[0])))
PY
