#!/bin/bash
# parity tests + C2 bench line + the c3/c4-shaped kernel launch lists (tail behaviour on repeat-rich data)
TAG=$1; BP=${2:-400000000}; PAIRS=${3:-200000}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log | head -1
timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
python - gpurun_out/${TAG}_c2.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); s = d["stage_ms_last_step"]
print(f"C2 {d['value']/1e6:7.3f} M/s  e2e {d['e2e']['value']/1e6:7.3f}  " + " ".join(f"{k[3:]} {v:.2f}" for k, v in s.items()))
PY
export LARIAT_C3_BP=$BP
KERN='regex:smem_kernel|sa_lookup_kernel|chain_kernel|extend_kernel|extend_heavy_kernel|matesw_kernel|pair_prepare_kernel|global_cigar_kernel|cigar_plan_kernel|cigar_band_kernel|seed_finish_kernel|smem_redo_kernel|pack_kernel'
for CFG in c3 c4; do
  timeout 900 python bench.py --config $CFG --pairs $PAIRS --steps 2 --warmup 2 --no-cpu > gpurun_out/${TAG}_${CFG}.json 2> gpurun_out/${TAG}_${CFG}.err
  python - gpurun_out/${TAG}_${CFG}.json $CFG <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); s = d["stage_ms_last_step"]
    print(f"{sys.argv[2]} {d['value']/1e3:7.1f} k/s  e2e {d['e2e']['value']/1e3:7.1f}  " + " ".join(f"{k[3:]} {v:.1f}" for k, v in s.items()))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 80 --csv --log-file gpurun_out/${TAG}_${CFG}_launches.csv python bench.py --config $CFG --pairs $PAIRS --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_${CFG}_ncu.log 2>&1
done
