python -m pytest tests/test_fft_gpu.py -m gpu -x -q 2>&1 | tail -5
python tools/fft_timing.py 16 20 24 2>&1 | tail -3
CURVE=1 python tools/fft_timing.py 22 2>&1 | tail -1
