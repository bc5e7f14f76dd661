#!/bin/bash
# generalized slice rule: sizes on both curves (compare with profiles/r02_window_sweep.log / r2_run26), then the full suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run41_slices.log
: > $out
echo "== BN254" >> $out; PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 20 21 22 23 24 2>&1 | grep log_L | cut -c40-170 >> $out
echo "== BN254 2^22 forced 8 waves (S=90)" >> $out; MSM_B200_SLICE=90 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c40-170 >> $out
echo "== BLS" >> $out; CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 20 21 22 2>&1 | grep log_L | cut -c40-170 >> $out
echo "== batched" >> $out; CHUNKS=1024 PRECOMPUTE_CHUNKED=1 timeout 120 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c40-190 >> $out
cat $out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -m gpu -x -q > gpurun_out/r2_run41_pytest.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2_run41_pytest.log
