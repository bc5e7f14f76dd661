./tools/dfma_mul gpurun_out/dfma_samples.txt > gpurun_out/dfma_mul.json; cat gpurun_out/dfma_mul.json
python tools/jobs/check_dfma.py gpurun_out/dfma_samples.txt
