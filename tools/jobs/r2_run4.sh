#!/bin/bash
# Round 2, GPU job 4: affine rounds with L2 prefetch + lazy planes.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_fullsize_golden.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for r in 0 1 2 3 4 5; do
  echo "ROUNDS=$r"; MSM_B200_BA_ROUNDS=$r PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1
done
echo "BPS=3 ROUNDS=4"; MSM_B200_BA_BPS=3 MSM_B200_BA_ROUNDS=4 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 24 2>&1 | tail -1
echo "== per-kernel"
PRECOMPUTE=0 MSM_B200_BA_ROUNDS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_affine_round|k_accumulate' -c 20 --csv --log-file gpurun_out/launches_r02_d.csv python tools/quick_timing.py 24 > gpurun_out/ncu_r02_d.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_r02_d.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum': print(d['Kernel Name'][:58], d['Metric Value'])
PY
echo "== BLS12-381 2^22"; for r in 0 3; do CURVE=1 MSM_B200_BA_ROUNDS=$r PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 22 2>&1 | tail -1; done
