#!/bin/bash
# c3 / c4 shaped workloads at the largest reference the on-GPU index builder manages (LARIAT_C3_BP), 32- and 64-bit interval arithmetic.
# usage: gpurun --timeout 1500 -- bash scripts/gpu_c34.sh <tag> <bp> <pairs>
TAG=$1; BP=${2:-400000000}; PAIRS=${3:-1000000}
mkdir -p gpurun_out
export LARIAT_C3_BP=$BP
for CFG in c3 c4; do
  ( time timeout 1200 python bench.py --config $CFG --pairs $PAIRS --steps 3 --warmup 3 --cpu-sample 50000 ) > gpurun_out/${TAG}_${CFG}.json 2> gpurun_out/${TAG}_${CFG}.err
  tail -4 gpurun_out/${TAG}_${CFG}.err; cut -c1-200 gpurun_out/${TAG}_${CFG}.json
done
LC_FORCE_WIDE=1 timeout 600 python bench.py --config c3 --pairs $PAIRS --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_c3_wide.json 2> gpurun_out/${TAG}_c3_wide.err
cut -c1-200 gpurun_out/${TAG}_c3_wide.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
