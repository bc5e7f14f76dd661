#!/bin/bash
# Build "o" on 8 GPUs: in-library multi-device path (tests + timing), NCCL bench at N = 2, 4, 8, configs[3] and [4] at N = 8.
# EVERY multi-rank command runs under its own `timeout`.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T="timeout 200"
echo "== multi-device tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device or sharded_resident" -rs 2>&1 | tail -4
for n in 8 4 2; do
  echo "== bench N=$n"
  $T python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_r02_o_n$n.json 2> gpurun_out/bench_r02_o_n$n.err || echo "N=$n failed or timed out"
done
echo "== batched N=8"; $T python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload batched --steps 5 --warmup 3 > gpurun_out/bench_r02_o_batched_n8.json 2> gpurun_out/bench_r02_o_batched_n8.err || echo "batched N=8 failed or timed out"
echo "== BLS N=8"; $T python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --curve 1 --log-n 22 --steps 5 --warmup 3 > gpurun_out/bench_r02_o_bls_n8.json 2> gpurun_out/bench_r02_o_bls_n8.err || echo "BLS N=8 failed or timed out"
for f in n2 n4 n8 batched_n8 bls_n8; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_o_$f.json'))
    print('$f', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('config', {}).get('window_bits'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'), 'agree', d.get('paths_agree'), 'golden', d.get('result_matches_golden'), d.get('phases_ms'))
except Exception as e:
    print('$f', 'no result:', e)
PY
done
