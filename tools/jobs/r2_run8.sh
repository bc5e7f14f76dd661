#!/bin/bash
# Round 2, GPU job 8 (1 GPU): geometric sub-batches (e2e), NTT with shared-memory twiddles, full parity.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== e2e 2^24: default (4 geometric), ratio sweep, 8 equal"
timeout 300 python - <<'PY'
import os, subprocess, sys
for lg, cfgs in ((24, [("", ""), ("", "1"), ("", "3"), ("8", "1"), ("3", "2"), ("5", "2")]), (21, [("", ""), ("2", "1"), ("2", "3"), ("3", "2")]), (22, [("", ""), ("2", "1"), ("3", "2")])):
    for depth, ratio in cfgs:
        env = dict(os.environ)
        if ratio: env["MSM_B200_PIPELINE_RATIO"] = ratio
        args = [sys.executable, "tools/e2e_timing.py", str(lg)] + ([depth] if depth else ["0"])
        out = subprocess.run(args, env=env, capture_output=True, text=True).stdout.strip().splitlines()
        print("log_n", lg, "depth", depth or "policy", "ratio", ratio or "2(default)", out[-1] if out else "no output", flush=True)
PY
echo "== scalar FFT"; timeout 300 python tools/fft_timing.py 2>&1 | tail -6
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_h_n1.json 2> gpurun_out/bench_r02_h_n1.err; tail -2 gpurun_out/bench_r02_h_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_h_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['phases_ms'], d['result_matches_golden'], d['no_table']['ms_per_step'])"
