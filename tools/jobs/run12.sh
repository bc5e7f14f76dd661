MSM_B200_SORT=binned PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | grep log_L
for m in atomic binned; do echo $m; MSM_B200_SORT=$m PRECOMPUTE=0 python tools/quick_timing.py 14 16 18 20 2>&1 | grep log_L | cut -c1-175; done
for m in atomic binned; do echo $m; MSM_B200_SORT=$m python tools/quick_timing.py 14 16 18 20 22 2>&1 | grep log_L | cut -c1-175; done
for m in atomic binned; do echo $m; MSM_B200_SORT=$m CHUNKS=1024 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c1-175;  done
for m in atomic binned; do echo $m; MSM_B200_SORT=$m LINES=10 CHUNKS=2048 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 21 2>&1 | grep log_L | cut -c1-175;  done
