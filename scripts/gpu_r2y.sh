#!/bin/bash
# chunk plans of the e2e arm: size of the first chunk
TAG=${1:-r2y}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk" ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu --e2e-sweep 500000,500000:LC_CHUNK_FIRST=60,500000:LC_CHUNK_FIRST=50,500000:LC_CHUNK_FIRST=40,500000:LC_CHUNK_FIRST=30,500000:LC_CHUNK_FIRST=20,400000:LC_CHUNK_FIRST=50,500000:LC_CHUNK_FIRST=50+LC_LANES=3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
grep "e2e sweep" gpurun_out/${TAG}_bench_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'])"
for F in 50 35; do
LC_CHUNK_FIRST=$F timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c2_first$F.json 2> gpurun_out/${TAG}_bench_c2_first$F.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2_first$F.json')); print('C2 first $F', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'])"
done
