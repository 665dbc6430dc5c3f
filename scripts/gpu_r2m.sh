#!/bin/bash
# matesw_kernel capped at 128 registers, small batches with warp-per-candidate CIGARs, e2e chunk sizes on C2, the C4 shape at full reference size
TAG=${1:-r2m}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --config c3 --pairs 200000 --steps 2 --warmup 2 --cpu-sample 20000 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
tail -1 gpurun_out/${TAG}_c3.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c3.json')); print('C3', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
timeout 900 python bench.py --config c2 --steps 3 --warmup 3 --e2e-sweep 200000,250000,333334,500000 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
grep -i "sweep\|chunk" gpurun_out/${TAG}_c2.err | head
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e'].get('pageable'), j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
timeout 600 python bench.py --mode percall --threads 32,128,512 --steps 4 --no-cpu > gpurun_out/${TAG}_percall.json 2> gpurun_out/${TAG}_percall.err
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_percall.json'))
for p in j['percall']: print('percall', p['threads'], round(p['pairs_per_s']), p['call_ms_p50'], p['call_ms_p99'], p.get('batch_device_ms_mean'))"
timeout 1200 python bench.py --config c4 --pairs 200000 --steps 2 --warmup 1 --cpu-sample 10000 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
tail -2 gpurun_out/${TAG}_c4.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c4.json')); print('C4', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'], j['cpu_baseline'])"
