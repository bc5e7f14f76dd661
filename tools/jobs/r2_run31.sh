#!/bin/bash
# window model with the reduction's latency floor: sizes 2^18 .. 2^22 on both curves (table path and plain path),
# and a sweep of forced window sizes on small plain calls
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run31_windows.log
: > $out
echo "== BN254 table path" >> $out;  PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 18 19 20 21 22 2>&1 | grep log_L | cut -c40-190 >> $out
echo "== BN254 plain path" >> $out;  MSM_B200_TABLE=off timeout 300 python tools/quick_timing.py 16 18 19 20 21 2>&1 | grep log_L | cut -c40-190 >> $out
echo "== BLS table path" >> $out;  CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 18 19 20 22 2>&1 | grep log_L | cut -c40-190 >> $out
echo "== BLS plain path" >> $out;  CURVE=1 MSM_B200_TABLE=off timeout 300 python tools/quick_timing.py 16 18 19 20 2>&1 | grep log_L | cut -c40-190 >> $out
echo "== BN254 plain, forced windows, 2^14 / 2^16 / 2^18" >> $out; WINDOWS=0,8,10,12,13,14,15,16 MSM_B200_TABLE=off timeout 300 python tools/quick_timing.py 14 16 18 2>&1 | grep log_L | cut -c40-190 >> $out
echo "== BN254 table, forced table windows at 2^20" >> $out
for c in 17 18 19 20; do PRECOMPUTE=$c timeout 100 python tools/quick_timing.py 20 2>&1 | grep log_L | cut -c40-190 >> $out; done
echo "== BLS table, forced table windows at 2^19" >> $out
for c in 16 17 18 19 20; do CURVE=1 PRECOMPUTE=$c timeout 100 python tools/quick_timing.py 19 2>&1 | grep log_L | cut -c40-190 >> $out; done
cat $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_run31_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_run31_pytest.log
