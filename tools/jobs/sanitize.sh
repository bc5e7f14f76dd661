export MSM_B200_SORT=binned
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned_sort_path and 13-1-1-1 or pipelined_sub_batches and 3-12 or window_table_chunked and 2-64-64" 2>&1 | tail -8
echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned_sort_path and 0-13-1-1-1" 2>&1 | tail -8
echo "racecheck rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ec_fft_gpu.py -m gpu -x -q -k "vs_oracle and 0-5" 2>&1 | tail -4
