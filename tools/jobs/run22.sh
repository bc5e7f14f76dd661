for q in 16 32 64 128 256; do echo Q=$q; MSM_B200_REDUCE_Q=$q LINES=10 CHUNKS=2048 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 21 2>&1 | grep log_L | cut -c60-200; done
for q in 8 16 32 64 128; do echo Q=$q; MSM_B200_REDUCE_Q=$q CHUNKS=1024 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c60-200; done
