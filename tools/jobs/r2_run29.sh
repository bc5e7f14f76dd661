#!/bin/bash
# ncu --set full of the two big sort kernels at 2^24 (table c = 22)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PRECOMPUTE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_partition|k_bin_place|k_bin_count|k_bin_hist" --launch-skip 8 --launch-count 4 -o gpurun_out/r02_sort_kernels python tools/quick_timing.py 24 > gpurun_out/r2_run29_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_sort_kernels.ncu-rep
