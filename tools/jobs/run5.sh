export MSM_B200_PIPELINE_DEVICE=1
for d in 1 2 4 8; do MSM_B200_PIPELINE=$d PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | tail -1; done
