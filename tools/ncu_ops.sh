#!/bin/bash
# ncu --set full of selected kernels of the op-level benchmark -> gpurun_out/<tag>.ncu-rep
# usage: tools/ncu_ops.sh <tag> <kernel regex> <skip> <count>
tag=$1; regex=$2; skip=$3; count=$4
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count \
    -f -o gpurun_out/$tag python tools/op_bench.py ncu > gpurun_out/$tag.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/$tag.log; ls -la gpurun_out/$tag.ncu-rep
