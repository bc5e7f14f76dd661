#!/bin/bash
# BLS12-381 2^22: cost of the 2-way split of a device-resident row, and where it goes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
out=gpurun_out/r2_run39_bls_split.log
: > $out
for d in 1 2; do for ov in 0 1; do
  echo "== split=$d overlap=$ov" >> $out
  CURVE=1 MSM_B200_SORT_OVERLAP=$ov MSM_B200_PIPELINE_RATIO=3 MSM_B200_PIPELINE_DEVICE=1 MSM_B200_PIPELINE=$d PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c40-170 >> $out
done; done
cat $out
for d in 1 2; do
CURVE=1 MSM_B200_SORT_OVERLAP=0 MSM_B200_PIPELINE_RATIO=3 MSM_B200_PIPELINE_DEVICE=1 MSM_B200_PIPELINE=$d PRECOMPUTE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2_run39_launches_split$d.csv python tools/quick_timing.py 22 > gpurun_out/r2_run39_ncu$d.log 2>&1
done
