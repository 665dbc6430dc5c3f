#!/bin/bash
# resident CTAs of the latency-bound chain kernel
TAG=${1:-r2ab}
mkdir -p gpurun_out
run() {  # config pairs name env...
  CFG=$1; PP=$2; N=$3; shift 3
  env "$@" timeout 600 python bench.py --config $CFG --pairs $PP --steps 3 --warmup 2 --no-cpu 2> gpurun_out/${TAG}_${CFG}_$N.err > gpurun_out/${TAG}_${CFG}_$N.json
  python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_${CFG}_$N.json')); s=j['stage_ms_last_step']; print('$CFG $N', round(j['value']), round(j['e2e']['value']), {k:round(v,2) for k,v in s.items()})"
}
run c2 1000000 chain4 LC_CHAIN_CTAS=4
run c2 1000000 chain6 LC_CHAIN_CTAS=6
run c2 1000000 chain7 LC_CHAIN_CTAS=7
run c2 1000000 chain3 LC_CHAIN_CTAS=3
