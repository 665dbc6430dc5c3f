export LARIAT_C3_BP=400000000
KERN='regex:smem_kernel|sa_lookup_kernel|chain_kernel|extend_kernel|matesw_kernel|pair_prepare_kernel|global_cigar_kernel|cigar_plan_kernel|cigar_band_kernel|seed_finish_kernel|smem_redo_kernel|pack_kernel'
python bench.py --config c3 --pairs 200000 --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 60 --csv --log-file gpurun_out/r01x_c3_launches.csv python bench.py --config c3 --pairs 200000 --steps 1 --warmup 1 --no-cpu > gpurun_out/r01x_c3_ncu.log 2>&1
tail -2 gpurun_out/r01x_c3_ncu.log | cut -c1-200
