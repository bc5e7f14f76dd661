#!/bin/bash
# BLS12-381: sort overlap off by rule (3 blocks per SM); slice lengths of fewer, just-under-whole waves
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run40_bls.log
: > $out
echo "== BLS e2e 2^22 default / overlap forced on" >> $out
CURVE=1 timeout 200 python tools/e2e_timing.py 22 0 2>&1 | grep e2e_ms | cut -c1-150 >> $out
CURVE=1 MSM_B200_SORT_OVERLAP=1 timeout 200 python tools/e2e_timing.py 22 0 2>&1 | grep e2e_ms | cut -c1-150 >> $out
echo "== BN254 e2e 2^24 default" >> $out
timeout 200 python tools/e2e_timing.py 24 0 2>&1 | grep e2e_ms | cut -c1-150 >> $out
echo "== BLS 2^22 slice sweep (11 10 9 8 7 6 5 4 waves)" >> $out
for s in 88 96 107 120 138 160 192 240; do
  echo "S=$s" >> $out
  CURVE=1 MSM_B200_SLICE=$s PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c40-170 >> $out
done
echo "== BLS 2^19 slice sweep" >> $out
for s in 0 24 32 40 48 64; do
  echo "S=$s" >> $out
  if [ $s = 0 ]; then CURVE=1 PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 19 2>&1 | grep log_L | cut -c40-170 >> $out
  else CURVE=1 MSM_B200_SLICE=$s PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 19 2>&1 | grep log_L | cut -c40-170 >> $out; fi
done
cat $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or sub_batch or task_groups or montgomery" 2>&1 | tail -2
