#!/bin/bash
# Parity tests, then the bench line for a list of environment settings (one per argument, e.g. "LC_SEED_GATE_MASK=7 LC_SEED_GATE_N=16").
# usage: gpurun --timeout 1200 -- bash scripts/gpu_sweep.sh <tag> "<env settings 1>" "<env settings 2>" ...
TAG=$1; shift
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
i=0
for SET in "$@"; do
  i=$((i+1))
  env $SET timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/${TAG}_sweep_$i.json 2> gpurun_out/${TAG}_sweep_$i.err
  python - "$SET" gpurun_out/${TAG}_sweep_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    s = d["stage_ms_last_step"]
    print(f"{sys.argv[1]:50s} {d['value']/1e6:7.3f} M/s  e2e {d['e2e']['value']/1e6:7.3f}  seed {s['ms_seed']:.2f} chain {s['ms_chain']:.2f} ext {s['ms_extend']:.2f} pair {s['ms_pair']:.2f} cigar {s['ms_cigar']:.2f} total {s['ms_total']:.2f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
