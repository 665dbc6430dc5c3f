#!/bin/bash
# warp-per-read chaining kernel: parity, then the C3 step with it on (default) and off
TAG=${1:-r2d}
P=${2:-200000}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
for CH in 192 0 48; do
LC_CHAIN_HEAVY=$CH timeout 900 python bench.py --config c3 --pairs $P --steps 2 --warmup 2 --cpu-sample 20000 > gpurun_out/${TAG}_c3_ch$CH.json 2> gpurun_out/${TAG}_c3_ch$CH.err
tail -2 gpurun_out/${TAG}_c3_ch$CH.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_c3_ch$CH.json')); print('chain_heavy=$CH', j['value'], j['stage_ms_last_step'], j['verified_pairs'])"
done
