MSM_B200_SORT=binned PRECOMPUTE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_binned.csv python tools/quick_timing.py 24 > gpurun_out/ncu_binned.log 2>&1
tail -1 gpurun_out/ncu_binned.log
