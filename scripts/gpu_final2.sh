#!/bin/bash
# End-of-round measurement (second half of round 2): parity, the driver's own invocation, the 3.1 Gb configs, the per-call arm, C1,
# ncu --set full facts for C2 and for the kernels of C3 that changed
TAG=${1:-r2zz}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'], j['cpu_baseline']['value'], j['roofline']['frac'], j['clocks'])"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_bench_c2_reference.json 2> gpurun_out/${TAG}_bench_c2_reference.err
cut -c1-200 gpurun_out/${TAG}_bench_c2_reference.json
for CFG in c3 c4; do
timeout 1200 python bench.py --config $CFG --pairs 200000 --steps 3 --warmup 3 --cpu-sample 20000 > gpurun_out/${TAG}_bench_$CFG.json 2> gpurun_out/${TAG}_bench_$CFG.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_$CFG.json')); print('$CFG', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'], j['cpu_baseline']['value'], j['roofline']['frac'])"
done
timeout 600 python bench.py --mode percall --threads 32,128,512 --steps 4 > gpurun_out/${TAG}_bench_percall_c2.json 2> gpurun_out/${TAG}_bench_percall_c2.err
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_bench_percall_c2.json'))
for p in j['percall']: print('percall', p['threads'], round(p['pairs_per_s']), p['call_ms_p50'], p['call_ms_p99'], p.get('batch_device_ms_mean'))"
timeout 900 python bench.py --config c1 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c1.json')); print('C1', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['cpu_baseline']['value'])"
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${TAG}_c2 \
    python bench.py --config c2 --pairs 1000000 --profile-pass > gpurun_out/${TAG}_c2_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_c2_ncu.log
ncu -i /tmp/${TAG}_c2.ncu-rep --page raw --csv > gpurun_out/${TAG}_c2_raw.csv 2> /dev/null
python scripts/ncu_kernels.py gpurun_out/${TAG}_c2_raw.csv gpurun_out/${TAG}_c2_kernels.json c2 1000000 "${TAG}_c2.ncu-rep (ncu --set full --clock-control none, profile pass of bench.py --config c2 --pairs 1000000)"
ncu -i /tmp/${TAG}_c2.ncu-rep --page details > gpurun_out/${TAG}_c2_details.txt 2> /dev/null
gzip -f gpurun_out/${TAG}_c2_raw.csv
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"cigar_band|cigar_plan|seed_finish|chain_kernel|pair_prepare_kernel" -f -o /tmp/${TAG}_c3 \
    python bench.py --config c3 --pairs 200000 --profile-pass > gpurun_out/${TAG}_c3_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_c3_ncu.log
ncu -i /tmp/${TAG}_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_c3_raw.csv 2> /dev/null
python scripts/ncu_kernels.py gpurun_out/${TAG}_c3_raw.csv gpurun_out/${TAG}_c3_kernels_partial.json c3 200000 "${TAG}_c3.ncu-rep (ncu --set full --clock-control none -k regex:cigar_band|cigar_plan|seed_finish|chain_kernel|pair_prepare_kernel, profile pass of bench.py --config c3 --pairs 200000)"
ncu -i /tmp/${TAG}_c3.ncu-rep --page details > gpurun_out/${TAG}_c3_details_partial.txt 2> /dev/null
gzip -f gpurun_out/${TAG}_c3_raw.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
    python bench.py --config c3 --pairs 200000 --profile-pass > gpurun_out/${TAG}_ncu_launches_c3.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_c3.csv 2>/dev/null | grep -v "cub::" | head -30
ls -la gpurun_out | tail -5
