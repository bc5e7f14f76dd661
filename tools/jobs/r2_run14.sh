#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench batched / headline"
timeout 600 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_k_batched_n1.json 2> gpurun_out/bench_r02_k_batched_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_k_batched_n1.json')); print('batched', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['window_bits'], d['result_matches_golden'], d['roofline']['whole_step_frac'])"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_k_n1.json 2> gpurun_out/bench_r02_k_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_k_n1.json')); print('headline', d['ms_per_step'], d['e2e']['ms_per_step'], d['result_matches_golden'], d['roofline']['whole_step_frac'])"
echo "== next rows"; timeout 900 python tests/perf/bench_next_rows.py > gpurun_out/next_rows_r02.jsonl 2> gpurun_out/next_rows_r02.err; cut -c1-400 gpurun_out/next_rows_r02.jsonl; tail -2 gpurun_out/next_rows_r02.err
