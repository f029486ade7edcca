#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_model_gpu.py tests/test_ddp_gpu.py -x -q -m gpu 2>&1 | tail -2
for cap in 56 48 64; do
  B200MART_DW_MAXCTAS=$cap timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dw cap $cap:', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done
