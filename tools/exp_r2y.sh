#!/bin/bash
mkdir -p gpurun_out
echo "== ops / stacks with the fused backward off (tiled two-pass on every shape), each keys variant"
for kv in 0 1 2; do
B200MART_ATTN_KEYS_VARIANT=$kv B200MART_ATTN_FUSED_BWD=0 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_stack_gpu.py tests/test_visual_bert_gpu.py -x -q -m gpu 2>&1 | tail -2
done
echo "== real sizes"
timeout 900 python -m pytest tests/test_families_base_gpu.py tests/test_visual_bert_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
for c in flava vilt; do
  for kv in 0 1 2; do
    B200MART_ATTN_KEYS_VARIANT=$kv timeout 400 python bench.py --config $c --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_y_${c}_$kv.json 2> gpurun_out/bench_y_${c}_$kv.err
    python -c "
import json; d=json.load(open('gpurun_out/bench_y_${c}_$kv.json')); print('$c keys-variant=$kv', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms loss', d['final_loss'])" || tail -3 gpurun_out/bench_y_${c}_$kv.err
  done
done
for c in visualbert large; do
    timeout 400 python bench.py --config $c --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_y_${c}.json 2> gpurun_out/bench_y_${c}.err
    python -c "
import json; d=json.load(open('gpurun_out/bench_y_${c}.json')); print('$c', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms loss', d['final_loss'])" || tail -3 gpurun_out/bench_y_${c}.err
done
