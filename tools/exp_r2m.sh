#!/bin/bash
mkdir -p gpurun_out
bash tools/ncu_full.sh attnbwd_r2m "attn_bwd_fused" 14 3
bash tools/ncu_full.sh fusionbwd_r2m "fusion_bwd" 2 2
timeout 300 python tools/timeline.py r2m > gpurun_out/timeline_r2m.log 2>&1 && python tools/timeline_phases.py gpurun_out/timeline_r2m.json.gz > gpurun_out/timeline_phases_r2m.txt 2>&1; tail -45 gpurun_out/timeline_r2m.log | head -60
