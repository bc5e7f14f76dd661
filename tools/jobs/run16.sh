python tests/perf/bench_next_rows.py > gpurun_out/next_rows.jsonl 2> gpurun_out/next_rows.err; tail -3 gpurun_out/next_rows.err; cut -c1-900 gpurun_out/next_rows.jsonl
python bench.py --curve 1 --log-n 22 --steps 5 --warmup 3 > gpurun_out/bench_r01_g_bls_n1.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r01_g_bls_n1.json
python bench.py --log-n 20 --steps 10 --warmup 3 > gpurun_out/bench_r01_g_2p20_n1.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r01_g_2p20_n1.json
