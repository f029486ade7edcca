#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r2r.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_r2r.txt)"
grep -E "FAILED|Error|^E " gpurun_out/pytest_r2r.txt | head -20
for c in flava vilt; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${c}_r2r.json 2> gpurun_out/bench_${c}_r2r.err; echo "bench $c rc $?"; tail -4 gpurun_out/bench_${c}_r2r.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_${c}_r2r.json')); r=d['roofline']
print('$c:', round(d['value'],1), 'ex/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],3), 'loss', d['final_loss'], 'launches', d['gpu_launches'])"
done
