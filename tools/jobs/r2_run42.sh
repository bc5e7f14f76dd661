#!/bin/bash
# final build "p": bench lines on one GPU
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_r02_p_n1.json 2> gpurun_out/bench_r02_p_n1.err || echo "bench failed"
timeout 300 python bench.py --curve 1 --log-n 22 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_p_bls_n1.json 2> gpurun_out/bench_r02_p_bls_n1.err || echo "BLS failed"
timeout 300 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_p_batched_n1.json 2> gpurun_out/bench_r02_p_batched_n1.err || echo "batched failed"
for f in n1 bls_n1 batched_n1; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_p_$f.json'))
    print('$f', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('config', {}).get('window_bits'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'), 'agree', d.get('paths_agree'), 'golden', d.get('result_matches_golden'), 'frac', round(d['roofline']['frac'],4), round(d['roofline'].get('whole_step_frac'),4))
except Exception as e:
    print('$f', 'no result:', e)
PY
done
