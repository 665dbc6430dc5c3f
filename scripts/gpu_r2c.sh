#!/bin/bash
# Round 2, third GPU visit: tests (stage 6b, fixed-point rescue lists), C3 full size with the new build, the heavy-read threshold of
# the extension stage swept, source-level ncu of the two kernels with the longest tails on C3, per-call arm with device-time split.
TAG=${1:-r2c}
P=${2:-200000}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
( time LC_VERBOSE=1 timeout 1500 python bench.py --config c3 --pairs $P --steps 3 --warmup 3 --cpu-sample 20000 ) > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
tail -4 gpurun_out/${TAG}_c3.err; python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c3.json')); print('C3', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'])"
for TH in "128 32" "1000000 1000000"; do
set -- $TH
LC_EXT_HEAVY_SEEDS=$1 LC_EXT_HEAVY_CHAINS=$2 timeout 900 python bench.py --config c3 --pairs $P --steps 2 --warmup 2 --no-cpu > gpurun_out/${TAG}_c3_heavy$1.json 2> gpurun_out/${TAG}_c3_heavy$1.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c3_heavy$1.json')); print('heavy>$1/$2', j['value'], j['stage_ms_last_step'])"
done
for K in chain_kernel matesw_kernel; do
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:^$K -c 1 -f -o /tmp/${TAG}_$K \
    python bench.py --config c3 --pairs $P --profile-pass > gpurun_out/${TAG}_ncu_$K.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_$K.log
ncu -i /tmp/${TAG}_$K.ncu-rep --page source --csv 2> /dev/null | gzip > gpurun_out/${TAG}_${K}_source.csv.gz
ncu -i /tmp/${TAG}_$K.ncu-rep --page details > gpurun_out/${TAG}_${K}_details.txt 2> /dev/null
done
( time timeout 600 python bench.py --mode percall --threads 32,512 --steps 4 --no-cpu ) > gpurun_out/${TAG}_percall_c2.json 2> gpurun_out/${TAG}_percall_c2.err
cat gpurun_out/${TAG}_percall_c2.json
ls -la gpurun_out | tail -12
