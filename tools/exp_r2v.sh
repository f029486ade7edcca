#!/bin/bash
mkdir -p gpurun_out
for c in flava vilt visualbert; do
  timeout 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --timeline > gpurun_out/bench_tl_$c.json 2> gpurun_out/bench_tl_$c.err; echo "bench $c rc $?"
  head -45 gpurun_out/timeline_$c.txt | cut -c1-150
done
