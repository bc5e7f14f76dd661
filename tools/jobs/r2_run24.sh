#!/bin/bash
# (1) adaptive sub-batch growth: new test + e2e; (2) slice length vs wave quantisation of k_accumulate
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "sub_batch or full_size_default_path or pipelined or montgomery_scalars_fused" > gpurun_out/r2_run24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run24_pytest.log
tail -4 gpurun_out/r2_run24_pytest.log; grep "sub-batches per call" gpurun_out/r2_run24_pytest.log
out=gpurun_out/r2_run24_slice.log
: > $out
for s in 0 300 320 330 332 333 334 336 340 350 380 443 500 664; do
  echo "== S=$s" >> $out
  if [ $s = 0 ]; then PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 24 2>&1 | grep log_L >> $out
  else MSM_B200_SLICE=$s PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 24 2>&1 | grep log_L >> $out; fi
done
cat $out | python -c "
import sys, json
s=None
for l in sys.stdin:
    if l.startswith('=='): s=l.strip()
    else:
        d=json.loads(l); print(s, d['total_ms'], d['acc_ms'])
"
