python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export MSM_B200_PIPELINE_DEVICE=1
for d in 1 2 8; do MSM_B200_PIPELINE=$d PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | tail -1; done
unset MSM_B200_PIPELINE_DEVICE
python tools/e2e_timing.py 24 1 4 8 2>&1 | tail -3
CHUNKS=1024 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 22 2>&1 | tail -1
LINES=10 CHUNKS=2048 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 21 2>&1 | tail -1
python tools/quick_timing.py 20 21 2>&1 | tail -2
