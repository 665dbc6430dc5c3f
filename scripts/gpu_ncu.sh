#!/bin/bash
# Full ncu captures (--set full, source counters) of kernels inside a bench workload.
# usage: gpurun --timeout 900 -- bash scripts/gpu_ncu.sh <tag> "<kernel regexes>" "<bench args>" [env settings]
TAG=$1; KS=$2; ARGS=${3:---scale 1.0}; SET=$4
mkdir -p gpurun_out
env $SET python bench.py $ARGS --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1      # data + index cache
for K in $KS; do
env $SET timeout 800 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_${K} \
    python bench.py $ARGS --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full_${K}.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full_${K}.log
done
