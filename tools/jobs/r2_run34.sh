#!/bin/bash
# Final build "o" (sort under accumulate): whole GPU suite, smoke, bench lines, launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -m gpu -x -q > gpurun_out/r2_run34_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_run34_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench N=1"; timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_o_n1.json 2> gpurun_out/bench_r02_o_n1.err || echo "bench failed"
echo "== bench N=1 again, defaults"; timeout 400 python bench.py > gpurun_out/bench_r02_o_n1_b.json 2> gpurun_out/bench_r02_o_n1_b.err || echo "bench failed"
echo "== bench batched"; timeout 300 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_o_batched_n1.json 2> gpurun_out/bench_r02_o_batched_n1.err || echo "batched failed"
echo "== bench BLS"; timeout 300 python bench.py --curve 1 --log-n 22 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_o_bls_n1.json 2> gpurun_out/bench_r02_o_bls_n1.err || echo "BLS failed"
for f in n1 n1_b batched_n1 bls_n1; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_o_$f.json'))
    print('$f', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('config', {}).get('window_bits'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'), 'agree', d.get('paths_agree'), 'golden', d.get('result_matches_golden'), d.get('phases_ms'), 'frac', round(d['roofline']['frac'],4), round(d['roofline'].get('whole_step_frac'),4), d['roofline'].get('kernel_ms_steps'))
except Exception as e:
    print('$f', 'no result:', e)
PY
done
