#!/usr/bin/env bash
# mc128 (config 3) at the rank counts given:  gpurun --gpus 8 -- 'bash scripts/gpu_mc_scale.sh tag 8 4'
tag=$1; shift
for n in "$@"; do
  if [ "$n" -eq 1 ]; then python bench.py --workload mc128 --steps 5 --warmup 3 > gpurun_out/${tag}_mc128_n$n.json 2> gpurun_out/${tag}_mc128_n$n.err
  else PULPO_MC_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --workload mc128 --steps 5 --warmup 3 > gpurun_out/${tag}_mc128_n$n.json 2> gpurun_out/${tag}_mc128_n$n.err; fi
  grep "mc trace\] rank 0" gpurun_out/${tag}_mc128_n$n.err | tail -3
  python -c "import json; d=json.load(open('gpurun_out/${tag}_mc128_n$n.json')); print('mc128 n=$n', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), 'ms', round(d['samples_per_s'],1), 'samples/s', 'e2e', round(d['e2e']['value'],2))"
done
