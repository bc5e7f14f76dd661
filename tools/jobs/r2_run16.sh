#!/bin/bash
cd "$(dirname "$0")/../.."
echo "== fft tests"; timeout 600 python -m pytest tests/test_fft_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== fft timing"; timeout 300 python tools/fft_timing.py 12 16 20 22 24 2>&1 | tail -5
CURVE=1 timeout 300 python tools/fft_timing.py 20 24 2>&1 | tail -2
