python -m pytest tests/test_ec_fft_gpu.py -m gpu -x -q 2>&1 | tail -5
python tools/ecfft_timing.py 8 11 14 16 2>&1 | tail -4
CURVE=1 python tools/ecfft_timing.py 11 2>&1 | tail -1
