for c in 19 20 21; do PRECOMPUTE=$c python tools/quick_timing.py 21 2>&1 | grep log_L | cut -c1-190; done
for q in 8 16 32; do MSM_B200_REDUCE_Q=$q PRECOMPUTE=20 python tools/quick_timing.py 21 2>&1 | grep log_L | cut -c1-190; done
for c in 20 21; do PRECOMPUTE=$c python tools/quick_timing.py 22 2>&1 | grep log_L | cut -c1-190; done
for c in 21 22; do PRECOMPUTE=$c python tools/quick_timing.py 23 2>&1 | grep log_L | cut -c1-190; done
