python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 21; do PRECOMPUTE=$c python tools/quick_timing.py 21 2>&1 | grep log_L | cut -c1-190; done
for c in 0 17 18 19; do PRECOMPUTE=$c python tools/quick_timing.py 20 2>&1 | grep log_L | cut -c1-190; done
PRECOMPUTE=0 python tools/quick_timing.py 21 22 23 24 2>&1 | grep log_L | cut -c1-190
