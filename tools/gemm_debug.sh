#!/bin/bash
for d in 0 1 2 3 4 5 6 7; do
  echo "=== B200MART_GEMM_DEBUG=$d"
  B200MART_GEMM_DEBUG=$d timeout 300 python tools/gemm_bench.py --few 2>&1 | tail -n +2
done
