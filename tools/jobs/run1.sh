set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_table" 2>&1 | tail -5
CHUNKS=1024 python tools/quick_timing.py 22 2>&1 | tail -2
CHUNKS=1024 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 22 2>&1 | tail -3
LINES=10 CHUNKS=2048 python tools/quick_timing.py 21 2>&1 | tail -2
LINES=10 CHUNKS=2048 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 21 2>&1 | tail -3
