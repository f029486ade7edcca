#!/bin/bash
mkdir -p gpurun_out
B200MART_GEMM_EPI16=-1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/pytest_r2e.txt 2>&1; echo "pytest(epi16) rc $? : $(tail -1 gpurun_out/pytest_r2e.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2e.txt | head -10
bash tools/ab_env.sh "X=0" "B200MART_GEMM_EPI16=-1" "B200MART_GEMM_EPI16=12" "B200MART_GEMM_EPI16=44" "B200MART_GEMM_EPI16=62" "X=1" "B200MART_GEMM_EPI16=-1" 2>&1 | tee gpurun_out/ab_r2e.txt
B200MART_GEMM_EPI16=-1 timeout 300 python tools/gemm_bench.py --few > gpurun_out/gemm_shapes_epi16_r2e.txt 2>&1; tail -14 gpurun_out/gemm_shapes_epi16_r2e.txt | cut -c1-40,88-110
