#!/bin/bash
mkdir -p gpurun_out
for kc in 32 16; do
B200MART_ATTN_KC=$kc timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_stack_gpu.py tests/test_base_gpu.py -x -q -m gpu > gpurun_out/pytest_r2n_$kc.txt 2>&1; echo "pytest kc=$kc rc $? : $(tail -1 gpurun_out/pytest_r2n_$kc.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2n_$kc.txt | head -10
done
bash tools/ab_env.sh "B200MART_ATTN_KC=64" "B200MART_ATTN_KC=32" "B200MART_ATTN_KC=16" "B200MART_ATTN_KC=64" "B200MART_ATTN_KC=32" "B200MART_ATTN_KC=16" 2>&1 | tee gpurun_out/ab_r2n.txt
for kc in 64 32 16; do echo "== KC $kc"; B200MART_PDL=0 B200MART_ATTN_KC=$kc timeout 300 python tools/op_bench.py kc$kc 2>/dev/null | grep attn_bwd; done
