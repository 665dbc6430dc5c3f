#!/bin/bash
# B-tree nodes with chain positions; occupancy knobs of the latency-bound kernels swept on C3 and C2
TAG=${1:-r2k}
P=${2:-200000}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
run() {  # name, env...
  N=$1; shift
  env "$@" timeout 600 python bench.py --config $CFG --pairs $PP --steps 2 --warmup 1 --no-cpu 2> gpurun_out/${TAG}_$N.err > gpurun_out/${TAG}_$N.json
  python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_$N.json')); s=j['stage_ms_last_step']; print('$CFG $N', round(j['value']), {k:round(v,1) for k,v in s.items()})"
}
CFG=c3; PP=$P
run base LC_X=0
run ch12 LC_CHAIN_HEAVY_CTAS=12 LC_CHAIN_CTAS=6
run ch12gl13 LC_CHAIN_HEAVY_CTAS=12 LC_CHAIN_CTAS=6 LC_SEED_GL_CTAS=13
run ch12h48 LC_CHAIN_HEAVY_CTAS=12 LC_CHAIN_CTAS=6 LC_CHAIN_HEAVY=48
run ch8 LC_CHAIN_HEAVY_CTAS=8 LC_CHAIN_CTAS=4
CFG=c2; PP=1000000
run base LC_X=0
run ch6 LC_CHAIN_CTAS=6 LC_CHAIN_HEAVY_CTAS=12
