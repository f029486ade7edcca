#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/pytest_r2d.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_r2d.txt)"
bash tools/ab_env.sh "X=0" "B200MART_GEMM_BEARLY=0" "X=1" "B200MART_GEMM_BEARLY=0" 2>&1 | tee gpurun_out/ab_r2d.txt
# ncu --set full with source: 40 consecutive pair-GEMM launches in the forward pass (skip the first 30), and 40 in the backward
bash tools/ncu_full.sh gemm_fwd_r2d gemm_bf16_tcgen05_pair 24 24
bash tools/ncu_full.sh attn_r2d "attn_(fwd_mma|bwd_fused)" 14 6
bash tools/ncu_full.sh fusion_r2d "fusion_" 0 3
bash tools/ncu_full.sh ln_r2d "layernorm_bwd" 10 3
