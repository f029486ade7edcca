#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_stack_gpu.py tests/test_base_gpu.py -x -q -m gpu > gpurun_out/pytest_r2p.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_r2p.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2p.txt | head -10
bash tools/ab_env.sh "X=0" "X=1" 2>&1 | tee gpurun_out/ab_r2p.txt
B200MART_PDL=0 timeout 300 python tools/op_bench.py r2p 2>/dev/null | grep attn_fwd
