MSM_B200_FIELD=u29 python -m pytest tests/test_gpu_parity.py tests/test_ec_fft_gpu.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
MSM_B200_FIELD=sat32 python -m pytest tests/test_gpu_parity.py tests/test_ec_fft_gpu.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
