#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py -x -q -m gpu -s 2>&1 | tail -8
for mode in eager graph; do
  flag=""; [ $mode = graph ] && flag="--graph"
  for rep in 1 2; do
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu $flag > gpurun_out/bench_g_${mode}_$rep.json 2> gpurun_out/bench_g_${mode}_$rep.err; echo "bench $mode rc $?"
    python -c "
import json; d=json.load(open('gpurun_out/bench_g_${mode}_$rep.json')); print('$mode', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('pixels_from_host'), d['gpu_launches'], d['final_loss'])" || tail -5 gpurun_out/bench_g_${mode}_$rep.err
  done
done
