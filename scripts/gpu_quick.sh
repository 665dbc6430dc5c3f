#!/bin/bash
# parity tests, the C2 line, the c3/c4-shaped lines (400 Mb, 1M pairs) — stage times only
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/quick_pytest.log 2>&1; tail -1 gpurun_out/quick_pytest.log
timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_last_step']; print('C2', round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), {k[3:]:round(v,2) for k,v in s.items()})"
export LARIAT_C3_BP=400000000
for CFG in c3 c4; do
timeout 900 python bench.py --config $CFG --pairs ${1:-1000000} --no-cpu --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_last_step']; print('$CFG', round(d['value']/1e3,1), round(d['e2e']['value']/1e3,1), d['config'].get('resident_batch_pairs'), {k[3:]:round(v,1) for k,v in s.items()})"
done
