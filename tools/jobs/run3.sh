python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | tail -2
CURVE=1 PRECOMPUTE=0 python tools/quick_timing.py 22 2>&1 | tail -1
CHUNKS=1024 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 22 2>&1 | tail -1
