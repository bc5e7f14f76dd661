#!/bin/bash
cd "$(dirname "$0")/../.."
for q in 4 8 16 32; do
  echo "REDUCE_Q=$q"
  MSM_B200_REDUCE_Q=$q PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 21 24 2>&1 | grep log_L | cut -c40-200
  MSM_B200_REDUCE_Q=$q CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 2>&1 | grep log_L | cut -c40-200
done
