#!/bin/bash
# source-level ncu of the two kernels that carry the C2 step
TAG=${1:-r2n}
KS=${2:-"extend_kernel smem_kernel"}
mkdir -p gpurun_out
python bench.py --config c2 --profile-pass > /dev/null 2> gpurun_out/${TAG}_prep.err
for K in $KS; do
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:^$K" -c 1 -f -o /tmp/${TAG}_$K \
    python bench.py --config c2 --profile-pass > gpurun_out/${TAG}_ncu_$K.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_$K.log
ncu -i /tmp/${TAG}_$K.ncu-rep --page source --csv 2> /dev/null | gzip > gpurun_out/${TAG}_c2_${K}_source.csv.gz
ncu -i /tmp/${TAG}_$K.ncu-rep --page details > gpurun_out/${TAG}_c2_${K}_details.txt 2> /dev/null
done
ls -la gpurun_out | tail -6
