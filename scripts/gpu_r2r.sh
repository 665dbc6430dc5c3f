#!/bin/bash
# staged copies of lc_align_packed (upload turns, regions copied out under the CIGAR kernels): parity, then the e2e arm per chunk size
TAG=${1:-r2r}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --e2e-sweep 200000,250000,333334,400000,500000,1000000 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
grep "e2e sweep" gpurun_out/${TAG}_bench_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
timeout 900 python bench.py --config c3 --pairs 200000 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c3.json')); print('C3', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'])"
