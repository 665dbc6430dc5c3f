#!/bin/bash
# 80-cell CIGAR class: parity + C3 step; source-level ncu of the kernels that carry the C3 step
TAG=${1:-r2g}
P=${2:-200000}
KS=${3:-"extend_heavy_kernel matesw_kernel cigar_band_kernel chain_heavy_kernel"}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --config c3 --pairs $P --steps 2 --warmup 2 --cpu-sample 20000 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
tail -2 gpurun_out/${TAG}_c3.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c3.json')); print('C3', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
for K in $KS; do
F=$(echo $K | tr -d '<>')
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:^$K" -c 1 -f -o /tmp/${TAG}_$F \
    python bench.py --config c3 --pairs $P --profile-pass > gpurun_out/${TAG}_ncu_$F.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_$F.log
ncu -i /tmp/${TAG}_$F.ncu-rep --page source --csv 2> /dev/null | gzip > gpurun_out/${TAG}_${F}_source.csv.gz
ncu -i /tmp/${TAG}_$F.ncu-rep --page details > gpurun_out/${TAG}_${F}_details.txt 2> /dev/null
done
ls -la gpurun_out | tail -12
