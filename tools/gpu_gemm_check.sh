#!/bin/bash
# first GPU bring-up of the tcgen05 GEMM: per-layout probes in separate processes (a trap poisons the context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for w in nt nn tn tk; do
  timeout 120 python tools/gemm_probe.py $w > gpurun_out/probe_$w.txt 2>&1
  echo "exit $?" >> gpurun_out/probe_$w.txt
done
for t in test_gemm_nt test_gemm_nn_b_mn test_gemm_tn_both_mn test_gemm_a_mn_only test_gemm_epilogue_fwd test_gemm_epilogue_dact; do
  timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k $t -x 2>&1 | tail -15 > gpurun_out/pytest_$t.txt
done
tail -n 30 gpurun_out/probe_*.txt
tail -n 5 gpurun_out/pytest_*.txt
