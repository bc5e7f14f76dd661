#!/bin/bash
# Round 2, GPU job 7 (1 GPU): full parity suite on the build with the dedicated squaring, the single-decomposition
# partition with packed pairs, the slice floor; timings; bench line; ncu launch list + --set full of k_accumulate.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== timings"
PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 20 22 24 2>&1 | grep log_L
CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 22 2>&1 | grep log_L
echo "no table:"; MSM_B200_TABLE=off timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1
echo "== bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_g_n1.json 2> gpurun_out/bench_r02_g_n1.err; tail -2 gpurun_out/bench_r02_g_n1.err; cut -c1-300 gpurun_out/bench_r02_g_n1.json
echo "== ncu launch list (bench, 1 step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_g.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_g.log 2>&1; tail -1 gpurun_out/ncu_r02_g.log | cut -c1-120
echo "== ncu --set full k_accumulate (device-resident 2^24, table)"
PRECOMPUTE=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_accumulate' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_accumulate -f python tools/quick_timing.py 24 > gpurun_out/ncu_r02_g2.log 2>&1; tail -1 gpurun_out/ncu_r02_g2.log | cut -c1-160
ls -la gpurun_out/r02_accumulate.ncu-rep
echo "== e2e pipeline depth, shards of 2^20 .. 2^22"
for lg in 20 21 22; do timeout 300 python tools/e2e_timing.py $lg 1 2 4 8 2>&1 | grep pipeline; done
