#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the same command and one full capture
# of the dominant kernel.  Usage (from the build container):  gpurun --timeout 1500 -- bash scripts/gpu_round.sh [tag] [scale]
TAG=${1:-r01}
SCALE=${2:-1.0}
KERN='regex:smem_kernel|sa_lookup_kernel|chain_kernel|extend_kernel|extend_heavy_kernel|matesw_kernel|pair_prepare_kernel|global_cigar_kernel|cigar_plan_kernel|cigar_band_kernel|seed_finish_kernel|smem_redo_kernel|pack_kernel|convert_kernel|DeviceScan'
TOP=${3:-smem_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
if ! grep -q " passed" gpurun_out/${TAG}_pytest.log || grep -q " failed" gpurun_out/${TAG}_pytest.log; then echo "parity tests not green: stopping"; exit 1; fi
( time timeout 900 python bench.py --scale $SCALE ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --scale $SCALE --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log
for K in $TOP; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K -s 2 -c 1 -f -o gpurun_out/${TAG}_${K} \
    python bench.py --scale $SCALE --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full_${K}.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full_${K}.log
done
ls -la gpurun_out
