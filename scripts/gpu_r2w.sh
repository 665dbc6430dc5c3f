#!/bin/bash
# end-of-row gate of the extension lane kernel: fixed period vs "as soon as N lanes wait"
TAG=${1:-r2w}
mkdir -p gpurun_out
run() {  # config pairs name env...
  CFG=$1; PP=$2; N=$3; shift 3
  env "$@" timeout 600 python bench.py --config $CFG --pairs $PP --steps 3 --warmup 2 --no-cpu 2> gpurun_out/${TAG}_${CFG}_$N.err > gpurun_out/${TAG}_${CFG}_$N.json
  python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_${CFG}_$N.json')); s=j['stage_ms_last_step']; print('$CFG $N', round(j['value']), {k:round(v,1) for k,v in s.items()})"
}
run c2 1000000 base LC_X=0
run c2 1000000 m3n8 LC_EXT_ROW_MASK=3 LC_EXT_ROW_N=8
run c2 1000000 m3n12 LC_EXT_ROW_MASK=3 LC_EXT_ROW_N=12
run c2 1000000 m7n10 LC_EXT_ROW_MASK=7 LC_EXT_ROW_N=10
run c2 1000000 m1n6 LC_EXT_ROW_MASK=1 LC_EXT_ROW_N=6
run c2 1000000 m3n16 LC_EXT_ROW_MASK=3 LC_EXT_ROW_N=16
run c3 200000 base LC_X=0
run c3 200000 m3n8 LC_EXT_ROW_MASK=3 LC_EXT_ROW_N=8
run c3 200000 m3n12 LC_EXT_ROW_MASK=3 LC_EXT_ROW_N=12
