python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-1200
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-300
LINES=10 CHUNKS=2048 PRECOMPUTE_CHUNKED=1 python tools/quick_timing.py 21 2>&1 | grep log_L | cut -c60-200
