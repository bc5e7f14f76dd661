export MSM_B200_PIPELINE_DEVICE=1
MSM_B200_PIPELINE=8 PRECOMPUTE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_sub8.csv python tools/quick_timing.py 24 > gpurun_out/ncu_sub8.log 2>&1
tail -2 gpurun_out/ncu_sub8.log
for p in 1 2 4; do MSM_B200_SCATTER_PASSES=$p MSM_B200_PIPELINE=8 PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | tail -1; done
for p in 1 2 4; do MSM_B200_SCATTER_PASSES=$p MSM_B200_PIPELINE=4 PRECOMPUTE=0 python tools/quick_timing.py 24 2>&1 | tail -1; done
