#!/bin/bash
# two GPUs, final kernels: the 3.1 Gb config with ONE read set sharded by barcode group (strong scaling) next to the same set on one GPU
TAG=${1:-r2zz}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --config c3 --pairs 400000 --scaling strong --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c3_strong_n2.json 2> gpurun_out/${TAG}_bench_c3_strong_n2.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c3_strong_n2.json')); print('C3 strong N=2', j['value'], j['e2e']['value'], j['stage_ms_last_step'], j['scaling'], j['shard'])"
timeout 600 python bench.py --gpus 1 --config c3 --pairs 400000 --scaling strong --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c3_strong_n1.json 2> gpurun_out/${TAG}_bench_c3_strong_n1.err
python -c "
import json,sys; j=json.load(open('gpurun_out/${TAG}_bench_c3_strong_n1.json')); print('C3 strong N=1', j['value'], j['e2e']['value'], j['stage_ms_last_step'])"
