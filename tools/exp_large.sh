#!/bin/bash
mkdir -p gpurun_out
bash tools/ncu_launches.sh large --large
python tools/summarize_launches.py gpurun_out/launches_large.csv --detail > gpurun_out/launches_large.txt 2>&1; head -24 gpurun_out/launches_large.txt
bash tools/ncu_full.sh gemm_full_large gemm_bf16_tcgen05_pair 48 24 --large
bash tools/ncu_full.sh attn_large "attn_(fwd_tiled|bwd_rows_tiled|bwd_keys_tiled)" 6 3 --large
