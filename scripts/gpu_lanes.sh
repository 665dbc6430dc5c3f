#!/bin/bash
# e2e arm vs number of lanes / chunk size, C2 and c3-like
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02c_pytest.log 2>&1; tail -1 gpurun_out/r02c_pytest.log
for SET in "LC_LANES=2" "LC_LANES=3 LC_CHUNK_PAIRS=334000" "LC_LANES=3 LC_CHUNK_PAIRS=250000" "LC_LANES=4 LC_CHUNK_PAIRS=250000"; do
env $SET timeout 600 python bench.py --no-cpu --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C2 $SET', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3))"
done
export LARIAT_C3_BP=400000000
for SET in "LC_LANES=2 LC_CHUNK_PAIRS=250000" "LC_LANES=4 LC_CHUNK_PAIRS=125000" "LC_LANES=4 LC_CHUNK_PAIRS=250000" "LC_LANES=6 LC_CHUNK_PAIRS=100000"; do
env $SET timeout 600 python bench.py --config c3 --pairs 1000000 --no-cpu --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 $SET', round(d['value']/1e3,1), 'e2e', round(d['e2e']['value']/1e3,1))"
done
