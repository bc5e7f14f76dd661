#!/bin/bash
# cost of the sub-batch split itself: device-resident scalars, forced split (no copy to wait for)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run22_split_cost.log
: > $out
for lg in 21 24; do
  for d in 1 2 3 4; do
    echo "== log_n=$lg split=$d" >> $out
    MSM_B200_PIPELINE_DEVICE=1 MSM_B200_PIPELINE=$d PRECOMPUTE=0 timeout 120 python tools/quick_timing.py $lg 2>&1 | grep log_L >> $out
  done
done
echo "== ratio 1 (equal halves), log_n=21 split=2" >> $out
MSM_B200_PIPELINE_RATIO=1 MSM_B200_PIPELINE_DEVICE=1 MSM_B200_PIPELINE=2 PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 21 2>&1 | grep log_L >> $out
for d in 1 2; do
MSM_B200_PIPELINE_DEVICE=1 MSM_B200_PIPELINE=$d PRECOMPUTE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2_run22_launches_split$d.csv python tools/quick_timing.py 21 > gpurun_out/r2_run22_ncu$d.log 2>&1
done
cat $out
