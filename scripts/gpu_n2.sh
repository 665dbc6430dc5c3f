#!/bin/bash
# two GPUs: the 3.1 Gb config with ONE read set sharded by barcode group (strong scaling), the C2 weak line, one process driving both GPUs
TAG=${1:-r2u}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 1500 $TR bench.py --gpus 2 --config c3 --pairs 400000 --scaling strong --steps 3 --warmup 3 --cpu-sample 20000 > gpurun_out/${TAG}_bench_c3_strong_n2.json 2> gpurun_out/${TAG}_bench_c3_strong_n2.err
tail -3 gpurun_out/${TAG}_bench_c3_strong_n2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c3_strong_n2.json')); print('C3 strong N=2', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['verified_pairs'], j['scaling'], j['shard'])"
timeout 900 python bench.py --gpus 1 --config c3 --pairs 400000 --scaling strong --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c3_strong_n1.json 2> gpurun_out/${TAG}_bench_c3_strong_n1.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c3_strong_n1.json')); print('C3 strong N=1', j['value'], j['e2e']['value'], j['stage_ms_last_step'])"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c2_n2.json 2> gpurun_out/${TAG}_bench_c2_n2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c2_n2.json')); print('C2 weak N=2', j['value'], j['e2e']['value'], j['stage_ms_last_step'])"
LC_PERCALL_DEVICES=2 timeout 600 python bench.py --mode percall --threads 64,256,1024 --steps 4 --no-cpu > gpurun_out/${TAG}_bench_percall_c2_oneproc_2gpu.json 2> gpurun_out/${TAG}_bench_percall_c2_oneproc_2gpu.err
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_bench_percall_c2_oneproc_2gpu.json'))
for p in j['percall']: print('percall 2 GPUs one process', p['threads'], round(p['pairs_per_s']), p['call_ms_p50'], p['call_ms_p99'])"
