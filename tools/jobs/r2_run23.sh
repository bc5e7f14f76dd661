#!/bin/bash
# e2e (host scalars) at 2^24 / 2^23 for several pipeline depths x growth ratios
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run23_e2e_ratio.log
: > $out
for lg in 24 23; do
for q in 2 2.5 3 3.4 4; do
  echo "== log_n=$lg ratio=$q" >> $out
  MSM_B200_PIPELINE_RATIO=$q timeout 200 python tools/e2e_timing.py $lg 2 3 4 5 2>&1 | grep e2e_ms >> $out
done
done
cat $out
