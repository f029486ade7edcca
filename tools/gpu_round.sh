#!/bin/bash
# One gpurun call that produces everything a tuning iteration needs (about 2 minutes of box time):
#   tools/gpu_round.sh <tag> [ENV=VALUE ...]
# 1. the GPU suite in ONE pytest process (the way the driver runs it), 2. bench.py (no CPU leg), 3. CUPTI step timeline + phases,
# 4. per-CTA GEMM phase stamps, 5. non-GEMM kernels alone (PDL off).  Outputs under gpurun_out/*_<tag>.*
tag=${1:-x}; shift
mkdir -p gpurun_out
for kv in "$@"; do export "$kv"; done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all_$tag.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_all_$tag.txt)"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    r = d["roofline"]
    print(f"{d['value']:.1f} ex/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.1f}  gemm {r['achieved']:.0f} TF/s ({r['frac']:.3f})  "
          f"attn {r['attention_ms_per_step']:.2f} fusion {r['fusion_ms_per_step']:.2f} ms  loss {d['final_loss']:.4f}")
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 300 python tools/timeline.py $tag > gpurun_out/timeline_$tag.log 2>&1 && python tools/timeline_phases.py gpurun_out/timeline_$tag.json.gz > gpurun_out/timeline_phases_$tag.txt 2>&1
head -12 gpurun_out/timeline_phases_$tag.txt
timeout 300 python tools/gemm_timeline.py $tag > gpurun_out/gemm_timeline_$tag.log 2>&1; tail -1 gpurun_out/gemm_timeline_$tag.log
B200MART_PDL=0 timeout 300 python tools/op_bench.py $tag > /dev/null 2>&1; cat gpurun_out/op_bench_$tag.txt
