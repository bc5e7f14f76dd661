set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k multi_device 2>&1 | tail -2
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_h_n1.json 2> gpurun_out/bench_r01_h_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_h_n$n.json 2> gpurun_out/bench_r01_h_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_h_batched_n8.json 2> gpurun_out/bench_r01_h_batched_n8.err
python bench.py --gpus 1 --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_h_batched_n1.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --curve 1 --log-n 22 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_h_bls_n8.json 2> gpurun_out/bench_r01_h_bls_n8.err
python bench.py --gpus 1 --curve 1 --log-n 22 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_h_bls_n1.json 2>/dev/null
for f in n1 n2 n4 n8 batched_n1 batched_n8 bls_n1 bls_n8; do python -c "
import json,sys
d=json.load(open('gpurun_out/bench_r01_h_$f.json'))
print('$f', '%.4g'%d['value'], '%.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], '%.3f'%d['e2e']['ms_per_step'], d['config']['window_bits'], d['phases_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['paths_agree'])
"; done
