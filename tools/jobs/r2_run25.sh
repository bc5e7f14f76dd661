#!/bin/bash
# wave-aligned slices: tests, then slice-length / reduce-Q sweeps at 2^21 and 2^24, then e2e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "sub_batch or full_size_default_path or pipelined or montgomery_scalars_fused" > gpurun_out/r2_run25_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run25_pytest.log
tail -3 gpurun_out/r2_run25_pytest.log; grep "sub-batches per call" gpurun_out/r2_run25_pytest.log
out=gpurun_out/r2_run25_sweep.log
: > $out
run() { echo "== $1" >> $out; shift; env "$@" PRECOMPUTE=0 timeout 120 python tools/quick_timing.py $LG 2>&1 | grep log_L >> $out; }
LG=24
run "2^24 default" X=1
for s in 381 444 533 666 888 1024; do run "2^24 S=$s" MSM_B200_SLICE=$s; done
LG=21
run "2^21 default" X=1
for s in 45 52 60 72 90 120 180; do run "2^21 S=$s" MSM_B200_SLICE=$s; done
for q in 8 16 32 64; do run "2^21 Q=$q" MSM_B200_REDUCE_Q=$q; done
LG=22
run "2^22 default" X=1
for q in 16 32 64; do run "2^22 Q=$q" MSM_B200_REDUCE_Q=$q; done
python - <<PY
import json
s=None
for l in open("$out"):
    if l.startswith('=='): s=l.strip()
    else:
        d=json.loads(l); print(s, 'total', d['total_ms'], 'sort', d['sort_ms'], 'acc', d['acc_ms'], 'red', d['red_ms'])
PY
echo "== e2e"
timeout 200 python tools/e2e_timing.py 24 0 2>&1 | grep e2e_ms
