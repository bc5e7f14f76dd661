#!/bin/bash
cd "$(dirname "$0")/../.."
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_golden.py -m gpu -x -q 2>&1 | tail -3
PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 21 22 24 2>&1 | grep log_L | cut -c40-200
CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 22 2>&1 | grep log_L | cut -c40-200
CHUNKS=1024 PRECOMPUTE_CHUNKED=1 timeout 300 python tools/quick_timing.py 22 2>&1 | tail -1 | cut -c40-200
echo "no table:"; MSM_B200_TABLE=off timeout 300 python tools/quick_timing.py 20 24 2>&1 | grep log_L | cut -c40-200
