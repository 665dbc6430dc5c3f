mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01z_pytest.log 2>&1; tail -1 gpurun_out/r01z_pytest.log
timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_last_step']; print('C2', round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), {k:round(v,2) for k,v in s.items()})"
export LARIAT_C3_BP=400000000
for SET in "LC_X=1" "LC_SEED_PK_CAP2=2048 LC_SEED_LST_CAP2=150" "LC_EXT_HEAVY_SEEDS=64 LC_EXT_HEAVY_CHAINS=16" "LC_EXT_HEAVY_SEEDS=12 LC_EXT_HEAVY_CHAINS=4"; do
env $SET timeout 600 python bench.py --config c3 --pairs 200000 --no-cpu --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_last_step']; print('c3 $SET', round(d['value']/1e3,1), {k:round(v,1) for k,v in s.items()})"
done
