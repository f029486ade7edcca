#!/bin/bash
mkdir -p gpurun_out
B200MART_ATTN_OCC4=1 B200MART_ATTN_FUSED_BWD=0 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_families_base_gpu.py tests/test_visual_bert_gpu.py -x -q -m gpu 2>&1 | tail -2
for c in vilt flava large; do
  for o in 0 1; do
    B200MART_ATTN_OCC4=$o timeout 400 python bench.py --config $c --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c occ4=$o', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms')"
  done
done
