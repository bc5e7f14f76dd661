#!/bin/bash
# bench stability with the 100 ms clock sampler: per-step kernel times of repeated runs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bn254 run $i: %.3f ms'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], d['roofline']['kernel_ms_steps'], d['clocks'])"
done
for i in 1 2 3; do
timeout 300 python bench.py --curve 1 --log-n 22 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_r02_q_bls_n1.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bls run $i: %.3f ms'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], d['roofline']['kernel_ms_steps'], d['clocks']['samples'], d['clocks']['sm_mhz'])"
done
