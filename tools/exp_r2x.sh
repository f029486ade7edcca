#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_families_base_gpu.py tests/test_visual_bert_gpu.py tests/test_model_gpu.py tests/test_base_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_.*_tiled -s 24 -c 2 -f -o gpurun_out/attn_tiled_vilt python bench.py --config vilt --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_attn_tiled.log 2>&1; echo "ncu rc $?"; ls -la gpurun_out/attn_tiled_vilt.ncu-rep
