#!/bin/bash
mkdir -p gpurun_out
echo "== real sizes (tiled forward + backward)"
timeout 900 python -m pytest tests/test_families_base_gpu.py tests/test_visual_bert_gpu.py tests/test_model_gpu.py tests/test_stack_gpu.py -x -q -m gpu -s 2>&1 | grep -i "passed\|failed\|rel err\|error" | tail -12
for c in flava vilt visualbert large; do
  for ft in 1 0; do
    B200MART_ATTN_FWD_TILED=$ft timeout 400 python bench.py --config $c --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_z_${c}_$ft.json 2> gpurun_out/bench_z_${c}_$ft.err
    python -c "
import json; d=json.load(open('gpurun_out/bench_z_${c}_$ft.json')); print('$c fwd-tiled=$ft', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms loss', d['final_loss'])" || tail -3 gpurun_out/bench_z_${c}_$ft.err
  done
done
