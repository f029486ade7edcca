#!/bin/bash
# One gpurun call: the measurements profiles/ cites for the current binary.
#   tools/final_r2.sh <tag>
# GPU suite (one pytest process, as the driver runs it), smoke(), bench.py on the four BASELINE configs + the reference arm,
# ncu launch list of one step, ncu --set full of 24 consecutive GEMM launches.
tag=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$tag.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_$tag.txt)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.txt 2>&1; echo "smoke rc $? : $(tail -1 gpurun_out/smoke_$tag.txt)"
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${tag}_base.json 2> gpurun_out/bench_${tag}_base.err; echo "bench base rc $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo "bench reference rc $?"
for c in large flava vilt visualbert; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err; echo "bench $c rc $?"
done
python - <<PY
import json
for c in ("base", "reference", "large", "flava", "vilt", "visualbert"):
    try:
        d = json.load(open("gpurun_out/bench_${tag}_%s.json" % c))
        r = d.get("roofline") or {}
        print(c, f"{d['value']:.1f} {d['unit']}  {d.get('ms_per_step', 0):.3f} ms  e2e {d['e2e']['value']:.1f}  frac {r.get('frac')}  clocks {d.get('clocks')}")
    except Exception as e:
        print(c, "unreadable:", e)
PY
bash tools/ncu_launches.sh $tag
python tools/summarize_launches.py gpurun_out/launches_$tag.csv --detail > gpurun_out/launches_$tag.txt 2>&1; head -30 gpurun_out/launches_$tag.txt
bash tools/ncu_full.sh gemm_full_$tag gemm_bf16_tcgen05_pair 24 24
