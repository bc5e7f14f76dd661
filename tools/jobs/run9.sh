python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned or two_level or pipelined" 2>&1 | tail -5
for m in atomic binned; do MSM_B200_SORT=$m PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | tail -1; done
for m in atomic binned; do MSM_B200_SORT=$m python tools/quick_timing.py 24 2>&1 | tail -1; done
for m in atomic binned; do MSM_B200_SORT=$m PRECOMPUTE=0 python tools/quick_timing.py 21 2>&1 | tail -1; done
for m in atomic binned; do MSM_B200_SORT=$m CURVE=1 PRECOMPUTE=0 python tools/quick_timing.py 22 2>&1 | tail -1; done
