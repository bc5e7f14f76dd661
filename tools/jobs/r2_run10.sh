#!/bin/bash
# Round 2, GPU job 10 (1 GPU): tree bucket reduction; parity; timings.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== timings (tree reduce on / off)"
for t in 1 0; do
  echo "REDUCE_TREE=$t"
  MSM_B200_REDUCE_TREE=$t PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 20 21 22 24 2>&1 | grep log_L | cut -c1-230
  MSM_B200_REDUCE_TREE=$t CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 22 2>&1 | grep log_L | cut -c1-230
done
echo "no table 2^24:"; MSM_B200_TABLE=off timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1 | cut -c1-230
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_i_n1.json 2> gpurun_out/bench_r02_i_n1.err; tail -2 gpurun_out/bench_r02_i_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_i_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['phases_ms'], d['result_matches_golden'], d['paths_agree'], d['no_table']['ms_per_step'])"
