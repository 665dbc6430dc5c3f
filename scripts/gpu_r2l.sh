#!/bin/bash
# Low-latency route for small batches: parity, the per-call arm (one barcode group per call from T threads) with and without it
TAG=${1:-r2l}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for SR in 4096 0 16384; do
LC_SMALL_READS=$SR timeout 600 python bench.py --mode percall --threads 32,128,512 --steps 4 --no-cpu > gpurun_out/${TAG}_percall_sr$SR.json 2> gpurun_out/${TAG}_percall_sr$SR.err
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_percall_sr$SR.json'))
for p in j['percall']: print('small_reads=$SR', p['threads'], round(p['pairs_per_s']), p['call_ms_p50'], p['call_ms_p99'], p.get('batch_device_ms_mean'))"
done
timeout 900 python bench.py --config c3 --pairs 200000 --steps 2 --warmup 2 --cpu-sample 20000 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
tail -1 gpurun_out/${TAG}_c3.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c3.json')); print('C3', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
