#!/bin/bash
# run each GPU test file in its own process (a device trap poisons the CUDA context) and keep the logs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for t in "$@"; do
  name=$(basename $t .py)
  timeout 600 python -m pytest $t -q -x -s -m gpu 2>&1 | tail -60 > gpurun_out/pytest_$name.txt
  s=${PIPESTATUS[0]}
  echo "exit $s" >> gpurun_out/pytest_$name.txt
  [ $s -ne 0 ] && rc=$s
  echo "=== $name (exit $s)"; tail -45 gpurun_out/pytest_$name.txt
done
exit $rc
