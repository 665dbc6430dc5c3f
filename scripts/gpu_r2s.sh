#!/bin/bash
# warp-aggregated pool atomics + tapered chunk plan: parity, stage times, the e2e arm per chunk plan and lane count
TAG=${1:-r2s}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --e2e-sweep 500000,500000:0,333334,250000,250000:0,1000000 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
grep "e2e sweep" gpurun_out/${TAG}_bench_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2.json')); print('C2', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
LC_LANES=3 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-sweep 500000,333334,250000,333334:0 > gpurun_out/${TAG}_bench_c2_lanes3.json 2> gpurun_out/${TAG}_bench_c2_lanes3.err
grep "e2e sweep" gpurun_out/${TAG}_bench_c2_lanes3.err
timeout 900 python bench.py --config c3 --pairs 200000 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c3.json')); print('C3', j['value'], j['e2e']['value'], j['e2e']['pageable']['value'], j['stage_ms_last_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_c2.csv \
    python bench.py --config c2 --profile-pass > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_c2.csv 2>/dev/null | grep -v "cub::" | head -30
