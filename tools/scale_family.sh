#!/bin/bash
# tools/scale_family.sh N config...: bench.py --config <c> at N GPUs (FLAVA / ViLT / VisualBERT arms: bucketed NCCL all-reduce after backward)
N=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config $c --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_${c}_n$N.json 2> gpurun_out/bench_${c}_n$N.err
  echo "bench $c N=$N rc $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_n$N.json')); print('$c N=$N', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'], d['clocks'])" || tail -5 gpurun_out/bench_${c}_n$N.err
done
