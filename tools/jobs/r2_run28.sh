#!/bin/bash
# ncu --set full of the BLS12-381 k_accumulate (2^22 points, table)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CURVE=1 PRECOMPUTE=0 timeout 200 python tools/quick_timing.py 22 2>&1 | grep log_L
CURVE=1 PRECOMPUTE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate --launch-skip 2 --launch-count 1 -o gpurun_out/r02_accumulate_bls python tools/quick_timing.py 22 > gpurun_out/r2_run28_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_accumulate_bls.ncu-rep
