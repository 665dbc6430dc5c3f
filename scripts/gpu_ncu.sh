#!/bin/bash
# One full ncu capture (--set full, source counters) of one kernel inside the bench workload.
# usage: gpurun --timeout 900 -- bash scripts/gpu_ncu.sh <tag> <kernel regex> [scale] [env settings]
TAG=$1; K=$2; SCALE=${3:-1.0}; SET=$4
mkdir -p gpurun_out
env $SET timeout 800 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/${TAG}_${K} \
    python bench.py --scale $SCALE --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full_${K}.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full_${K}.log
