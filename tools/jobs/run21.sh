python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned or full_size or pipelined" 2>&1 | tail -3
PRECOMPUTE=0 python tools/quick_timing.py 21 24 2>&1 | grep log_L | cut -c1-200
