#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_final3.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_final3.txt)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*_tiled -s 119 -c 3 -f -o gpurun_out/attn_tiled_vilt_final python bench.py --config vilt --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_attn_tiled_final.log 2>&1; echo "ncu rc $?"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('base', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
