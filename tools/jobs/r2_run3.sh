#!/bin/bash
# Round 2, GPU job 3: affine rounds with 3 / 4 resident blocks per SM, and an ncu --set full capture of one round.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for bps in 3 4; do for r in 2 4; do
  echo "BPS=$bps ROUNDS=$r"; MSM_B200_BA_BPS=$bps MSM_B200_BA_ROUNDS=$r PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1
done; done
echo "== per-kernel (ncu time only) BPS=4"
MSM_B200_BA_BPS=4 PRECOMPUTE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_affine_round|k_accumulate' -c 24 --csv --log-file gpurun_out/launches_r02_c.csv python tools/quick_timing.py 24 > gpurun_out/ncu_r02_c.log 2>&1
grep -o 'k_affine_round[^"]*\|k_accumulate[^"]*\|"[0-9,.]*"$' gpurun_out/launches_r02_c.csv | paste - - | tail -14 | cut -c1-60,200-
echo "== ncu --set full: one gather round and one plane round"
MSM_B200_BA_BPS=4 PRECOMPUTE=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_affine_round' --launch-skip 12 --launch-count 2 -o gpurun_out/r02_affine_round -f python tools/quick_timing.py 24 > gpurun_out/ncu_r02_c2.log 2>&1; tail -2 gpurun_out/ncu_r02_c2.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
