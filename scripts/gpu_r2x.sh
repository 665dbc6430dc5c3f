#!/bin/bash
# chunk plans of the e2e arm after the non-blocking region copy: equal chunks vs half-size first and last chunk
TAG=${1:-r2x}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk" ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu --e2e-sweep 500000,500000:1,333334:1,250000:1,250000,500000 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
grep "e2e sweep" gpurun_out/${TAG}_bench_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'])"
LC_CHUNK_TAPER=1 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c2_taper.json 2> gpurun_out/${TAG}_bench_c2_taper.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2_taper.json')); print('C2 taper', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'])"
