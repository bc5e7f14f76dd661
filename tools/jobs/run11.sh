python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned or two_level" 2>&1 | tail -2
MSM_B200_SORT=binned PRECOMPUTE=0 python tools/quick_timing.py 24 21 2>&1 | tail -2
