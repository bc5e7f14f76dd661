#!/bin/bash
# task-group pipelined upload of many-task rows: tests + batched bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_golden.py -x -q -m gpu -k "task_groups or batched or amt or pipelined or sub_batch or chunk" > gpurun_out/r2_run30_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_run30_pytest.log
timeout 300 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n_batched_n1.json 2> gpurun_out/bench_r02_n_batched_n1.err || echo "batched failed"
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_n_batched_n1.json'))
print('batched', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('config', {}).get('window_bits'), 'agree', d.get('paths_agree'), 'golden', d.get('result_matches_golden'))
PY
for g in 1 2 3 4 6 8; do
MSM_B200_PIPELINE=$g timeout 300 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('groups $g', 'e2e %.3f ms'%d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('paths_agree'), d.get('result_matches_golden'))"
done
