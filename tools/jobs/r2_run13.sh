#!/bin/bash
cd "$(dirname "$0")/../.."
for c in 10 11 12 13; do echo "batched 1024x4096 table c=$c"; CHUNKS=1024 PRECOMPUTE=$c timeout 300 python tools/quick_timing.py 22 2>&1 | tail -1 | cut -c40-230; done
for c in 9 10 11 12; do echo "AMT 10x2^21/2048 table c=$c"; LINES=10 CHUNKS=2048 PRECOMPUTE=$c timeout 300 python tools/quick_timing.py 21 2>&1 | tail -1 | cut -c40-230; done
for c in 12 13 14 15; do echo "AMT 10x2^21/128 table c=$c"; LINES=10 CHUNKS=128 PRECOMPUTE=$c timeout 300 python tools/quick_timing.py 21 2>&1 | tail -1 | cut -c40-230; done
for c in 16 17 18 19; do echo "2^20 table c=$c"; PRECOMPUTE=$c timeout 300 python tools/quick_timing.py 20 2>&1 | tail -1 | cut -c40-230; done
for c in 19 20 21 22; do echo "2^22 table c=$c"; PRECOMPUTE=$c timeout 300 python tools/quick_timing.py 22 2>&1 | tail -1 | cut -c40-230; done
for c in 19 20 21; do echo "2^21 table c=$c"; PRECOMPUTE=$c timeout 300 python tools/quick_timing.py 21 2>&1 | tail -1 | cut -c40-230; done
