#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r2g.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_r2g.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2g.txt | head -10
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_base_r2g.json 2> gpurun_out/bench_base_r2g.err; echo "bench base rc $?"; tail -3 gpurun_out/bench_base_r2g.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_base_r2g.json"))
r = d["roofline"]
print(f"base: {d['value']:.1f} ex/s {d['ms_per_step']:.3f} ms e2e {d['e2e']['value']:.1f} (h2d {d['e2e']['h2d_bytes_per_step']}) px-from-host {d['e2e']['pixels_from_host']['value']:.1f} gemm {r['achieved']:.0f} TF/s ({r['frac']:.3f}) step_frac {r['step_frac_of_peak']:.3f}")
print("eager", d["gpu_eager_baseline"]); print("cpu", d["cpu_baseline"])
PY
for c in large flava vilt; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${c}_r2g.json 2> gpurun_out/bench_${c}_r2g.err; echo "bench $c rc $?"; tail -4 gpurun_out/bench_${c}_r2g.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_${c}_r2g.json')); r=d['roofline']
print('$c:', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],3), 'loss', d['final_loss'], 'launches', d['gpu_launches'])"
done
