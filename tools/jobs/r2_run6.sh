#!/bin/bash
# Round 2, GPU job 6 (1 GPU): BLS12-381 small shards (the N = 8 anomaly), fused Montgomery scalars test.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "montgomery or affine_halving or edge or busy or abort" 2>&1 | tail -3
echo "== BLS12-381 2^19 .. 2^22 with table"
CURVE=1 PRECOMPUTE=0 timeout 300 python tools/quick_timing.py 19 20 21 22 2>&1 | grep "log_L\|precompute"
for c in 15 16 17 18 19; do echo "table c=$c"; CURVE=1 PRECOMPUTE=$c timeout 300 python tools/quick_timing.py 19 2>&1 | tail -1; done
echo "== per-kernel BLS 2^19"
CURVE=1 PRECOMPUTE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02_f.csv python tools/quick_timing.py 19 > gpurun_out/ncu_r02_f.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_r02_f.csv')))
hdr=None
out=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum': out.append((d['Kernel Name'][:58], d['Metric Value']))
for k,v in out[-26:]: print(k, v)
PY
