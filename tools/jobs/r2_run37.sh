#!/bin/bash
# buckets per reduction thread at 2^21 buckets (2^24 and 2^23 points, c = 22) and batched 1024 x 2^12
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run37_reduce_q.log
: > $out
for q in 0 16 32 64 128 256; do
  echo "== Q=$q (0 = default)" >> $out
  if [ $q = 0 ]; then PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 23 24 2>&1 | grep log_L | cut -c40-170 >> $out
  else MSM_B200_REDUCE_Q=$q PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 23 24 2>&1 | grep log_L | cut -c40-170 >> $out; fi
done
for q in 0 8 16 32 64; do
  echo "== batched Q=$q" >> $out
  if [ $q = 0 ]; then CHUNKS=1024 PRECOMPUTE_CHUNKED=1 timeout 120 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c40-190 >> $out
  else MSM_B200_REDUCE_Q=$q CHUNKS=1024 PRECOMPUTE_CHUNKED=1 timeout 120 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c40-190 >> $out; fi
done
cat $out
