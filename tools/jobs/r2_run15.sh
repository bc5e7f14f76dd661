#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio --clock-control none -k regex:'k_ntt' -c 12 --csv --log-file gpurun_out/ntt_r02.csv python tools/fft_timing.py 24 > gpurun_out/ntt_r02.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ntt_r02.csv')))
hdr=None
cur={}
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        print(d['ID'], d['Kernel Name'][:30], d['Metric Name'][:70], d['Metric Value'])
PY
