#!/bin/bash
TAG=${1:-r2ac}
mkdir -p gpurun_out
nproc
timeout 300 python bench.py --mode percall --threads 512,256 --steps 4 --no-cpu > gpurun_out/${TAG}_bench_percall_c2.json 2> gpurun_out/${TAG}_bench_percall_c2.err
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_bench_percall_c2.json'))
for p in j['percall']: print('percall', p['threads'], round(p['pairs_per_s']), p['call_ms_p50'], p['call_ms_p99'], p.get('batch_device_ms_mean'))"
