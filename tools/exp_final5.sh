#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_final5.txt 2>&1; echo "pytest rc $? : $(tail -1 gpurun_out/pytest_final5.txt)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_final5.json 2>gpurun_out/bench_final5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final5.json')); print('base', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config']['launch'][:40], d['gpu_launches'])"
