#!/bin/bash
# 16-bit two-per-register CIGAR band kernels: parity, C2/C3 lines, launch lists
TAG=${1:-r2v}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
timeout 900 python bench.py --config c3 --pairs 200000 --steps 3 --warmup 3 --cpu-sample 20000 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c3.json')); print('C3', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'], j['verified_pairs'])"
for CFG in "c2 1000000" "c3 200000"; do
set -- $CFG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_$1.csv \
    python bench.py --config $1 --pairs $2 --profile-pass > gpurun_out/${TAG}_ncu_launches_$1.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_$1.csv 2>/dev/null | grep "cigar"
done
