python -m pytest tests -m gpu -x -q -k "pipelined or window_table or edge or multiexp_vs_cpu" 2>&1 | tail -3
python tools/e2e_timing.py 24 1 2 4 8 2>&1 | tail -4
python tools/e2e_timing.py 22 1 2 4 8 2>&1 | tail -4
