#!/bin/bash
# 8-GPU probe: concurrent H2D bandwidth per rank (default vs NUMA-local pinned buffers) and the e2e call under both
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_run21_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/r2_run21_topo.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  tools/numa_h2d_probe.py 24 > gpurun_out/r2_run21_numa_probe.jsonl 2> gpurun_out/r2_run21_numa_probe.err || echo "probe failed or timed out"
tail -3 gpurun_out/r2_run21_numa_probe.err
wc -l gpurun_out/r2_run21_numa_probe.jsonl
