#!/bin/bash
# new full-size goldens: BLS12-381 batched, G2 at 2^20 / 2^18
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_golden.py -x -q -m gpu -k "bls381_batched or g2_fullsize" > gpurun_out/r2_run20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run20_pytest.log
tail -5 gpurun_out/r2_run20_pytest.log
