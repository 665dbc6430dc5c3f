#!/bin/bash
# Later seeding passes with their lists in global memory: parity, C3 / C2 steps with pass counts, launch list
TAG=${1:-r2i}
P=${2:-200000}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
LC_VERBOSE=1 timeout 900 python bench.py --config c3 --pairs $P --steps 2 --warmup 2 --cpu-sample 20000 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
grep "\[lc\]" gpurun_out/${TAG}_c3.err | sort | uniq -c | head; tail -2 gpurun_out/${TAG}_c3.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c3.json')); print('C3', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
timeout 600 python bench.py --config c2 --steps 3 --warmup 3 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c2.json')); print('C2', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['gpu_launches'])"
for V in "LC_SEED_LST_CAP2=40" "LC_SEED_LST_CAP2=96"; do
env $V LC_VERBOSE=1 timeout 600 python bench.py --config c3 --pairs $P --steps 2 --warmup 1 --no-cpu 2> gpurun_out/${TAG}_sweep.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$V', j['value'], j['stage_ms_last_step']['ms_seed'])"
grep "\[lc\]" gpurun_out/${TAG}_sweep.err | sort | uniq -c | head -4
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
    python bench.py --config c3 --pairs $P --profile-pass > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_c3.csv 2>/dev/null | head -12
