#!/bin/bash
# Round 2, second GPU visit: boundary tests, the per-call (drop-in) arm on C2, and ncu facts for every kernel on C2 and on the
# full-size C3 (profile pass only: one warm pass, one pass between cudaProfilerStart/Stop).
TAG=${1:-r2b}
C3PAIRS=${2:-200000}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py --mode percall --threads 32,128,512 --steps 4 ) > gpurun_out/${TAG}_percall_c2.json 2> gpurun_out/${TAG}_percall_c2.err
tail -3 gpurun_out/${TAG}_percall_c2.err; cat gpurun_out/${TAG}_percall_c2.json
for CFG in "c2 1000000" "c3 $C3PAIRS"; do
set -- $CFG
python bench.py --config $1 --pairs $2 --profile-pass > /dev/null 2> gpurun_out/${TAG}_$1_prep.err      # data + index cache
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${TAG}_$1 \
    python bench.py --config $1 --pairs $2 --profile-pass > gpurun_out/${TAG}_$1_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_$1_ncu.log
ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2> /dev/null
python scripts/ncu_kernels.py gpurun_out/${TAG}_$1_raw.csv gpurun_out/${TAG}_$1_kernels.json $1 $2 "${TAG}_$1.ncu-rep (ncu --set full, profile pass of bench.py --config $1 --pairs $2)"
ncu -i /tmp/${TAG}_$1.ncu-rep --page details > gpurun_out/${TAG}_$1_details.txt 2> /dev/null
gzip -f gpurun_out/${TAG}_$1_raw.csv
ls -la /tmp/${TAG}_$1.ncu-rep
done
ls -la gpurun_out | tail -20
