#!/bin/bash
# tools/scale_r2.sh N [timeline]: bench.py at N GPUs with the default settings (+ a CUPTI timeline of rank 0 for N > 1 when asked)
N=$1
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_scale_n$N.json 2> gpurun_out/bench_scale_n$N.err
fi
echo "bench N=$N rc $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_scale_n$N.json')); print('N=$N', round(d['value'],1), 'ex/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value'],1), 'replica', d.get('replica_check'), d['config'].get('gradient_exchange','')[:60], d['clocks'])" || tail -5 gpurun_out/bench_scale_n$N.err
if [ -n "$2" ] && [ "$N" != 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/timeline.py n$N > gpurun_out/timeline_n$N.log 2>&1; echo "timeline rc $?"
  python tools/timeline_phases.py gpurun_out/timeline_n$N.json.gz > gpurun_out/timeline_phases_n$N.txt 2>&1; head -40 gpurun_out/timeline_phases_n$N.txt
fi
