#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_stack_gpu.py -x -q -m gpu > gpurun_out/pytest_r2f.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_r2f.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2f.txt | head -10
bash tools/ab_env.sh "B200MART_LNB=1" "B200MART_LNB=2" "B200MART_LNB=1" "B200MART_LNB=2" 2>&1 | tee gpurun_out/ab_r2f.txt
B200MART_PDL=0 B200MART_LNB=1 timeout 300 python tools/op_bench.py lnb1 2>/dev/null | grep layernorm
B200MART_PDL=0 B200MART_LNB=2 timeout 300 python tools/op_bench.py lnb2 2>/dev/null | grep layernorm
