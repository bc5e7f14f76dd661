PRECOMPUTE=0 ncu --set full --clock-control none --import-source on -k regex:"k_partition|k_bin_place|k_bucket_reduce" -s 3 -c 3 -o gpurun_out/prof_sort_r01_h python tools/quick_timing.py 24 > gpurun_out/ncu_h.log 2>&1
tail -2 gpurun_out/ncu_h.log
