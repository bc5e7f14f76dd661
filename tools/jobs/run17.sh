python -m pytest tests -m gpu -x -q -k "concurrent or full_size_default" 2>&1 | tail -5
