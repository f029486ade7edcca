#!/bin/bash
# per-launch device time of one training step (cold-cache, serialised: compare SHARES) -> gpurun_out/launches_<tag>.csv
tag=${1:-r1}; shift
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py "$@" > gpurun_out/launches_$tag.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/launches_$tag.log; wc -l gpurun_out/launches_$tag.csv
