#!/bin/bash
# Round 2, first GPU job (1 GPU): full parity suite on the new build, bench line, affine-round sweeps.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T="timeout 900"
echo "== pytest -m gpu"; $T python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== quick timing 2^24 table: BA default / off / rounds sweep"
for r in default 0 1 2 3 4 5 6; do
  if [ $r = default ]; then unset MSM_B200_BA_ROUNDS; else export MSM_B200_BA_ROUNDS=$r; fi
  echo "BA_ROUNDS=$r"; PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1
done
unset MSM_B200_BA_ROUNDS
for b in 256 512 2048; do echo "BA_BATCH=$b"; MSM_B200_BA_BATCH=$b PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1; done
for g in 32 64 128; do echo "L2_FETCH=$g BA=0"; MSM_B200_BA=0 MSM_B200_L2_FETCH=$g PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1; done
for g in 32 128; do echo "L2_FETCH=$g BA on"; MSM_B200_L2_FETCH=$g PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1; done
echo "== other sizes/curves"
PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 20 22 2>&1 | tail -2
CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 22 2>&1 | tail -1
echo "no table 2^24:"; timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1
echo "== bench"
$T python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_a_n1.json 2> gpurun_out/bench_r02_a_n1.err; tail -3 gpurun_out/bench_r02_a_n1.err; cat gpurun_out/bench_r02_a_n1.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_a.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_a.log 2>&1; tail -2 gpurun_out/ncu_r02_a.log | cut -c1-300
