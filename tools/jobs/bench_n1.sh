set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_f_n1.json 2> gpurun_out/bench_r01_f_n1.err || exit 1
cat gpurun_out/bench_r01_f_n1.json | cut -c1-1500
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_f_ref.json 2>> gpurun_out/bench_r01_f_n1.err
cat gpurun_out/bench_r01_f_ref.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -o gpurun_out/prof_accumulate_r01_f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log
