#!/bin/bash
# End-of-round measurement: the driver's own invocation, the 3.1 Gb configs, the per-call arm, config 5, and ncu --set full facts per kernel
TAG=${1:-r2z}
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
tail -2 gpurun_out/${TAG}_bench_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'], j['cpu_baseline']['value'], j['clocks'])"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_bench_c2_reference.json 2> gpurun_out/${TAG}_bench_c2_reference.err
cat gpurun_out/${TAG}_bench_c2_reference.json | cut -c1-300
for CFG in c3 c4; do
timeout 1200 python bench.py --config $CFG --pairs 200000 --steps 3 --warmup 3 --cpu-sample 20000 > gpurun_out/${TAG}_bench_$CFG.json 2> gpurun_out/${TAG}_bench_$CFG.err
tail -1 gpurun_out/${TAG}_bench_$CFG.err
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
for CFG in "c2 1000000" "c3 200000"; do
set -- $CFG
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${TAG}_$1 \
    python bench.py --config $1 --pairs $2 --profile-pass > gpurun_out/${TAG}_$1_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_$1_ncu.log
ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2> /dev/null
python scripts/ncu_kernels.py gpurun_out/${TAG}_$1_raw.csv gpurun_out/${TAG}_$1_kernels.json $1 $2 "${TAG}_$1.ncu-rep (ncu --set full --clock-control none, profile pass of bench.py --config $1 --pairs $2)"
ncu -i /tmp/${TAG}_$1.ncu-rep --page details > gpurun_out/${TAG}_$1_details.txt 2> /dev/null
gzip -f gpurun_out/${TAG}_$1_raw.csv
done
( time timeout 900 python scripts/bench_micro.py ) > gpurun_out/${TAG}_config5_sweep.jsonl 2> gpurun_out/${TAG}_config5.err
tail -3 gpurun_out/${TAG}_config5.err; wc -l gpurun_out/${TAG}_config5_sweep.jsonl
ls -la gpurun_out | tail -14
