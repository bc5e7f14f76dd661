#!/bin/bash
# 2 vs 3 sub-batches for 2^22 / 2^21 scalars (both curves)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run38_small_pipeline.log
: > $out
for curve in 0 1; do for lg in 21 22; do
  echo "== curve=$curve log_n=$lg default" >> $out
  CURVE=$curve timeout 200 python tools/e2e_timing.py $lg 0 2>&1 | grep e2e_ms >> $out
  echo "== curve=$curve log_n=$lg 3 parts growth 3" >> $out
  CURVE=$curve MSM_B200_PIPELINE_RATIO=3 timeout 200 python tools/e2e_timing.py $lg 3 2>&1 | grep e2e_ms >> $out
  echo "== curve=$curve log_n=$lg 3 parts growth 2" >> $out
  CURVE=$curve MSM_B200_PIPELINE_RATIO=2 timeout 200 python tools/e2e_timing.py $lg 3 2>&1 | grep e2e_ms >> $out
done; done
cut -c1-140 $out
