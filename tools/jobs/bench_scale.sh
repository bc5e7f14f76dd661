set -x
nvidia-smi -L | head -8
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_g_n1.json 2> gpurun_out/bench_r01_g_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_g_n$n.json 2> gpurun_out/bench_r01_g_n$n.err
done
for n in 1 2 4 8; do python -c "
import json,sys
d=json.load(open('gpurun_out/bench_r01_g_n$n.json'))
print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['window_bits'], d['phases_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
"; done
