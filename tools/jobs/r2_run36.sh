#!/bin/bash
# sanity of the rebuilt library (msm_plan_describe added): plan tests on a device + a parity subset + one bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_plan_host_logic.py tests/test_abi.py -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_golden.py -m gpu -x -q -k "2pow24 or task_groups or sub_batch or pipelined or batched" 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1 %.3f ms'%d['ms_per_step'], 'e2e %.3f ms'%d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('paths_agree'), d.get('result_matches_golden'))"
python - <<'PY'
import ec_gpu_b200 as m
for curve, n, tc in ((0, 1 << 24, 22), (1, 1 << 22, 20), (2, 1 << 20, 0), (3, 1 << 18, 0)):
    p = m.describe_plan(curve, n, table_window_bits=tc)
    print(curve, n, {k: p[k] for k in ("window_bits", "num_windows", "slice_len", "slices", "wave_slices", "waves")})
PY
