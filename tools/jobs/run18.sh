python -m pytest tests/test_g2.py -m gpu -x -q 2>&1 | tail -8
CURVE=2 python tools/quick_timing.py 16 20 2>&1 | grep log_L | cut -c1-200
CURVE=3 python tools/quick_timing.py 16 20 2>&1 | grep log_L | cut -c1-200
CURVE=2 PRECOMPUTE=0 python tools/quick_timing.py 20 2>&1 | grep log_L | cut -c1-200
