python -m pytest tests/test_ec_fft_gpu.py tests/test_g2.py -m gpu -x -q 2>&1 | tail -3
MSM_B200_FIELD=u29 python -m pytest tests/test_ec_fft_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/ecfft_timing.py 8 11 14 16 2>&1 | tail -4
MSM_B200_ECFFT_GLV=0 python tools/ecfft_timing.py 11 16 2>&1 | tail -2
CURVE=1 python tools/ecfft_timing.py 11 2>&1 | tail -1
