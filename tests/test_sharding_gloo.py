"""The N > 1 path on the CPU: two gloo ranks shard one MSM the way bench.py / MultiexpKernel do
(contiguous ceil(n/N) chunks), compute their partials (with the oracle standing in for the GPU),
all-gather the 96-byte partial points and sum them; the result must equal the un-sharded MSM."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ec_gpu_b200 as m
    from oracle import oracle as O

    curve, seed = 0, 0x0BADC0DE
    start, end = m.shard_range(n, world, rank)
    pts = O.gen_points(curve, seed, end - start, start=start)
    sc = O.gen_scalars(curve, seed, end - start, start=start)
    partial = O.multiexp_cpu(curve, pts, sc, nthreads=1) if end > start else np.zeros(96, dtype=np.uint8)
    mine = torch.from_numpy(partial.copy())
    gathered = torch.zeros(world * 96, dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, mine)
    if rank == 0:
        acc = np.zeros((1, 96), dtype=np.uint8)
        for r in range(world):
            acc = O.ec_op(curve, 0, acc, gathered[r * 96:(r + 1) * 96].numpy().reshape(1, 96).copy())
        np.save(os.path.join(out_dir, "sum.npy"), acc)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_msm_equals_whole(tmp_path, oracle):
    n, world = 3001, 2  # odd size: shards of 1501 and 1500
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "sum.npy")
    pts, sc = oracle.gen_points(0, 0x0BADC0DE, n), oracle.gen_scalars(0, 0x0BADC0DE, n)
    want = oracle.multiexp_cpu(0, pts, sc)
    from util import assert_same_points

    assert_same_points(oracle, 0, got, want, "sharded")


def test_shard_ranges_cover_exactly(engine):
    for n in (0, 1, 7, 8, 9, 1 << 24, (1 << 24) + 5):
        for parts in (1, 2, 3, 4, 8):
            spans = [engine.shard_range(n, parts, i) for i in range(parts)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) == engine.chunk_size(n, parts) or n == 0
