#!/bin/bash
# round-2 experiment set A: baseline on this pool + GEMM tile switches + isolated shape sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r2a.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_r2a.txt)"
bash tools/ab_env.sh "X=0" "B200MART_GEMM_SMALLN=768" "B200MART_GEMM_SMALLN=3072" "B200MART_GEMM_PAIR=0" "B200MART_DW_MAXCTAS=0" "B200MART_DW_MAXCTAS=24" "B200MART_DW_MAXCTAS=64" "B200MART_GEMM_SMALLN=768 B200MART_DW_MAXCTAS=64" 2>&1 | tee gpurun_out/ab_r2a.txt
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_shapes_r2a.txt 2>&1; tail -30 gpurun_out/gemm_shapes_r2a.txt
