#!/bin/bash
# cooperative dedup in stage 5a, larger extension units first, occupancy of the warp seeding kernel x list capacity of the lane pass
TAG=${1:-r2q}
P=${2:-200000}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/${TAG}_pytest.log
run() {  # name, env...
  N=$1; shift
  env "$@" timeout 600 python bench.py --config $CFG --pairs $PP --steps 2 --warmup 1 --no-cpu 2> gpurun_out/${TAG}_$N.err > gpurun_out/${TAG}_$N.json
  python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_$N.json')); s=j['stage_ms_last_step']; print('$CFG $N', round(j['value']), {k:round(v,1) for k,v in s.items()})"
}
CFG=c3; PP=$P
run base LC_X=0
run redo8 LC_SEED_REDO_CTAS=8
run redo8cap48 LC_SEED_REDO_CTAS=8 LC_SEED_LST_CAP2=48
run redo8cap40 LC_SEED_REDO_CTAS=8 LC_SEED_LST_CAP2=40
run redo12cap32 LC_SEED_REDO_CTAS=12 LC_SEED_LST_CAP2=32
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
    python bench.py --config c3 --pairs $P --profile-pass > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_c3.csv 2>/dev/null | grep -v "cub::" | head -30
