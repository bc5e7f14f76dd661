#!/bin/bash
# Round 2, 8-GPU job: in-library multi-device path (tests + timing), NCCL bench at N = 1, 2, 4, 8, configs[3] and [4].
# EVERY multi-rank command runs under its own `timeout` (round 1 lost 117 GPU-minutes to one dead-locked torchrun).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1200 -- 'bash tools/jobs/r2_run5_8gpu.sh > gpurun_out/r2_run11.log 2>&1'
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T="timeout 200"
echo "== multi-device tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device or sharded_resident" -rs 2>&1 | tail -4
echo "== in-library multi-device timing"; timeout 400 python tools/multi_device_timing.py 24 > gpurun_out/multi_device_j.jsonl 2> gpurun_out/multi_device_j.err; cat gpurun_out/multi_device_j.jsonl; tail -2 gpurun_out/multi_device_j.err
echo "== bench N=1"; $T python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r02_j_n1.json 2> gpurun_out/bench_r02_j_n1.err || echo "N=1 failed"
for n in 2 4 8; do
  echo "== bench N=$n"
  $T python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_r02_j_n$n.json 2> gpurun_out/bench_r02_j_n$n.err || echo "N=$n failed or timed out"
done
echo "== batched N=8"; $T python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload batched --steps 5 --warmup 3 > gpurun_out/bench_r02_j_batched_n8.json 2> gpurun_out/bench_r02_j_batched_n8.err || echo "batched N=8 failed or timed out"
echo "== batched N=1"; $T python bench.py --gpus 1 --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_j_batched_n1.json 2> gpurun_out/bench_r02_j_batched_n1.err || echo "batched N=1 failed"
echo "== BLS N=1"; $T python bench.py --gpus 1 --curve 1 --log-n 22 --steps 5 --warmup 3 > gpurun_out/bench_r02_j_bls_n1.json 2> gpurun_out/bench_r02_j_bls_n1.err || echo "BLS N=1 failed"
echo "== BLS N=8"; $T python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --curve 1 --log-n 22 --steps 5 --warmup 3 > gpurun_out/bench_r02_j_bls_n8.json 2> gpurun_out/bench_r02_j_bls_n8.err || echo "BLS N=8 failed or timed out"
echo "== reference arm N=1 (CPU)"; timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_r02_j_ref.json 2> gpurun_out/bench_r02_j_ref.err || echo "ref failed"
for f in n1 n2 n4 n8 batched_n1 batched_n8 bls_n1 bls_n8 ref; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_j_$f.json'))
    print('$f', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d.get('config', {}).get('window_bits'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'), 'agree', d.get('paths_agree'), 'golden', d.get('result_matches_golden'), d.get('phases_ms'))
except Exception as e:
    print('$f', 'no result:', e)
PY
done
