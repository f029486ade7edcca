#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_base_gpu.py -x -q -m gpu > gpurun_out/pytest_r2q.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_r2q.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2q.txt | head -10
bash tools/ab_env.sh "X=0" "B200MART_ADAMW_CTAS=148" "B200MART_ADAMW_CTAS=296" "B200MART_ADAMW_CTAS=592" "B200MART_ADAMW_CTAS=74" "X=1" 2>&1 | tee gpurun_out/ab_r2q.txt
