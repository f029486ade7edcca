#!/bin/bash
# ncu --set full of selected kernels inside one profiled training step -> gpurun_out/<tag>.ncu-rep
# usage: tools/ncu_full.sh <tag> <kernel regex> <skip> <count> [profile_step args]
tag=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$regex -s $skip -c $count \
    -f -o gpurun_out/$tag python tools/profile_step.py "$@" > gpurun_out/$tag.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/$tag.log; ls -la gpurun_out/$tag.ncu-rep
