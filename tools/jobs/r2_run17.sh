#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for e in 1 0; do
  echo "REDUCE_EVEN=$e"
  MSM_B200_REDUCE_EVEN=$e PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 20 21 22 24 2>&1 | grep log_L | cut -c40-200
  MSM_B200_REDUCE_EVEN=$e CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 22 2>&1 | grep log_L | cut -c40-200
  echo "no table:"; MSM_B200_REDUCE_EVEN=$e MSM_B200_TABLE=off timeout 300 python tools/quick_timing.py 20 24 2>&1 | grep log_L | cut -c40-200
done
