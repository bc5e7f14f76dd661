for e in 12288 9216 7680 6144; do echo entries=$e; MSM_B200_PART_ENTRIES=$e PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | grep log_L | cut -c60-160; done
for e in 12288 7680; do echo entries=$e; MSM_B200_PART_ENTRIES=$e PRECOMPUTE=0 python tools/quick_timing.py 21 2>&1 | grep log_L | cut -c60-160; done
