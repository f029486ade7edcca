#!/bin/bash
mkdir -p gpurun_out
B200MART_GEMM_HALF=1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_stack_gpu.py -x -q -m gpu > gpurun_out/pytest_half_r2c.txt 2>&1; echo "pytest(half) rc $? : $(tail -1 gpurun_out/pytest_half_r2c.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_half_r2c.txt | head -20
bash tools/ab_env.sh "X=0" "B200MART_GEMM_HALF=1" "B200MART_GEMM_HALF=1 B200MART_GEMM_HALF_SLOTS=1" "B200MART_GEMM_HALF=1 B200MART_DW_MAXCTAS=64" "B200MART_GEMM_HALF=1 B200MART_DW_MAXCTAS=0" "B200MART_GEMM_HALF=1 B200MART_DW_MAXCTAS=96" 2>&1 | tee gpurun_out/ab_r2c.txt
B200MART_GEMM_HALF=1 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_shapes_half_r2c.txt 2>&1; tail -26 gpurun_out/gemm_shapes_half_r2c.txt | cut -c1-40,88-110
