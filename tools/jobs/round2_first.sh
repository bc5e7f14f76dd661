#!/bin/bash
# First GPU job of the next round (8-GPU box): everything the last 8-GPU job of round 1 could not finish.
# EVERY multi-rank command runs under its own `timeout`: a hung collective must cost minutes, not the round's
# whole GPU budget (round 1 lost 117 GPU-minutes to one dead-locked torchrun).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/jobs/round2_first.sh 2>&1 | tail -40'
T="timeout 150"
$T python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k multi_device 2>&1 | tail -2
$T python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_a_n1.json 2> gpurun_out/bench_r02_a_n1.err
for n in 2 4 8; do
  $T python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_a_n$n.json 2> gpurun_out/bench_r02_a_n$n.err || echo "N=$n failed or timed out"
done
$T python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_a_batched_n8.json 2> gpurun_out/bench_r02_a_batched_n8.err || echo "batched N=8 failed or timed out"
$T python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --curve 1 --log-n 22 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_a_bls_n8.json 2> gpurun_out/bench_r02_a_bls_n8.err || echo "BLS N=8 failed or timed out"
for f in n1 n2 n4 n8 batched_n8 bls_n8; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_a_$f.json'))
    print('$f', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['config']['window_bits'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['paths_agree'])
except Exception as e:
    print('$f', 'no result:', e)
PY
done
