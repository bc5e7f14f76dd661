#!/bin/bash
# Build "m" (wave-aligned slices, adaptive sub-batch growth): whole GPU suite, smoke, bench lines, launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -m gpu -x -q > gpurun_out/r2_run26_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_run26_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench N=1"; timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_m_n1.json 2> gpurun_out/bench_r02_m_n1.err || echo "bench failed"
echo "== bench batched"; timeout 300 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_m_batched_n1.json 2> gpurun_out/bench_r02_m_batched_n1.err || echo "batched failed"
echo "== bench BLS"; timeout 300 python bench.py --curve 1 --log-n 22 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_m_bls_n1.json 2> gpurun_out/bench_r02_m_bls_n1.err || echo "BLS failed"
echo "== bench 2^20"; timeout 300 python bench.py --log-n 20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_m_2p20_n1.json 2> gpurun_out/bench_r02_m_2p20_n1.err || echo "2^20 failed"
for f in n1 batched_n1 bls_n1 2p20_n1; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_m_$f.json'))
    print('$f', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('config', {}).get('window_bits'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'), 'agree', d.get('paths_agree'), 'golden', d.get('result_matches_golden'), d.get('phases_ms'), 'frac', d['roofline']['frac'], d.get('whole_step_frac'))
except Exception as e:
    print('$f', 'no result:', e)
PY
done
echo "== launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run26_ncu.log 2>&1; echo "ncu rc=$?"
echo "== other sizes"
PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 20 21 22 23 2>&1 | grep log_L | cut -c40-200
CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 22 2>&1 | grep log_L | cut -c40-200
for lg in 21 22 23; do timeout 200 python tools/e2e_timing.py $lg 0 2>&1 | grep e2e_ms; done
