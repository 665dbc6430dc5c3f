#!/bin/bash
# 3.1 Gb config: does cutting the 200 k-pair call into chunks over 2-3 lanes (latency-bound kernels of one chunk under the compute-bound ones of the other) help?
TAG=${1:-r2aa}
mkdir -p gpurun_out
timeout 900 python bench.py --config c3 --pairs 200000 --steps 3 --warmup 2 --no-cpu --e2e-sweep 0,100000,100000:LC_CHUNK_FIRST=100,67000,50000 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
grep "e2e sweep" gpurun_out/${TAG}_bench_c3.err
LC_LANES=3 timeout 900 python bench.py --config c3 --pairs 200000 --steps 3 --warmup 2 --no-cpu --e2e-sweep 67000,50000,40000 > gpurun_out/${TAG}_bench_c3_lanes3.json 2> gpurun_out/${TAG}_bench_c3_lanes3.err
grep "e2e sweep" gpurun_out/${TAG}_bench_c3_lanes3.err
