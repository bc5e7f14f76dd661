#!/bin/bash
# Round 2, GPU job 2 (1 GPU): affine rounds v2 (warp-contiguous mapping, planes, block-shared inversion).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== parity (golden full-size + small sizes)"; timeout 900 python -m pytest tests/test_fullsize_golden.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
echo "== 2^24 table: BA rounds sweep"
for r in 0 1 2 3 4 5 default; do
  if [ $r = default ]; then unset MSM_B200_BA_ROUNDS; else export MSM_B200_BA_ROUNDS=$r; fi
  echo "BA_ROUNDS=$r"; PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1
done
unset MSM_B200_BA_ROUNDS
for b in 128 256 512; do echo "BA_BATCH=$b"; MSM_B200_BA_BATCH=$b PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1; done
echo "== per-kernel times (ncu, serialised) of one device-resident 2^24 call"
PRECOMPUTE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_affine_round|k_accumulate|k_scan|k_halve|k_fixup|k_bucket_reduce|k_bin|k_partition' -c 120 --csv --log-file gpurun_out/launches_r02_b.csv python tools/quick_timing.py 24 > gpurun_out/ncu_r02_b.log 2>&1; tail -1 gpurun_out/ncu_r02_b.log | cut -c1-200
echo "== other sizes / curves"
PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 20 21 22 2>&1 | grep log_L
CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 22 2>&1 | tail -1
CURVE=1 MSM_B200_BA=0 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 22 2>&1 | tail -1
echo "== reference kernel on this GPU"
timeout 900 python tests/perf/ref_kernel_b200.py > gpurun_out/ref_kernel_b200.jsonl 2> gpurun_out/ref_kernel_b200.err; cat gpurun_out/ref_kernel_b200.jsonl; tail -3 gpurun_out/ref_kernel_b200.err
