#!/bin/bash
# Round 2, first GPU visit: parity tests (incl. the CUDA index builder), the full-size 3.1 Gb index + C3 line with run-time
# verification against the CPU reference, md5 of the index files (compared with `bwa index` run in the build container), C2 line.
TAG=${1:-r2a}
C3PAIRS=${2:-200000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt; free -g >> gpurun_out/${TAG}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
( time LC_VERBOSE=1 timeout 1500 python bench.py --config c3 --pairs $C3PAIRS --steps 3 --warmup 3 --cpu-sample 20000 ) > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
tail -12 gpurun_out/${TAG}_c3.err; cat gpurun_out/${TAG}_c3.json
md5sum data/cache/c3_*/c3.fa.bwt data/cache/c3_*/c3.fa.sa data/cache/c3_*/c3.fa.pac data/cache/c3_*/c3.fa.ann data/cache/c3_*/c3.fa.amb > gpurun_out/${TAG}_c3_md5.txt 2>&1
cat gpurun_out/${TAG}_c3_md5.txt
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
tail -4 gpurun_out/${TAG}_c2.err; cat gpurun_out/${TAG}_c2.json
