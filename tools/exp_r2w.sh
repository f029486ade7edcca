#!/bin/bash
mkdir -p gpurun_out
echo "== ops with the fused backward off (tiled two-pass on every op shape)"
B200MART_ATTN_FUSED_BWD=0 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_stack_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== families / visual bert / large at real sizes"
timeout 900 python -m pytest tests/test_families_base_gpu.py tests/test_visual_bert_gpu.py tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12
for c in flava vilt visualbert large; do
  for tiled in 1 0; do
    B200MART_ATTN_TILED=$tiled timeout 400 python bench.py --config $c --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_w_${c}_$tiled.json 2> gpurun_out/bench_w_${c}_$tiled.err
    python -c "
import json; d=json.load(open('gpurun_out/bench_w_${c}_$tiled.json')); print('$c tiled=$tiled', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms loss', d['final_loss'])" || tail -3 gpurun_out/bench_w_${c}_$tiled.err
  done
done
