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


def _timed_loop_worker(rank, world, port, out_dir):
    """The collective structure of bench.py's timed(): K steps that each hold an all-gather, the
    all-reduced time, then the extra steps of the clock sampler -- with rank-dependent step times."""
    import time

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    calls = 0
    mine, gathered = torch.full((4,), rank, dtype=torch.uint8), torch.zeros(4 * world, dtype=torch.uint8)

    def step():
        nonlocal calls
        time.sleep(0.001 * (1 + 3 * rank))  # ranks run at different speeds
        dist.all_gather_into_tensor(gathered, mine)
        calls += 1

    steps = 5
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    extra = bench.sampler_extra_steps(float(ms.item()), steps)
    for _ in range(min(extra, 40)):  # same clamp on every rank
        step()
    dist.barrier()
    np.save(os.path.join(out_dir, "calls%d.npy" % rank), np.array([calls, extra]))
    dist.destroy_process_group()


def test_bench_extra_steps_are_rank_uniform(tmp_path):
    """bench.py keeps the load up after a short timed region so that the clock sampler sees it; the step
    contains a collective, so the number of extra steps must be the same on every rank (a count taken
    from a local clock deadlocks NCCL at N = 8)."""
    import bench

    assert bench.sampler_extra_steps(300.0, 5) == 0
    assert bench.sampler_extra_steps(28.0, 5) == 54      # N = 8: 5 steps of 5.6 ms
    assert bench.sampler_extra_steps(0.0, 5) == 2000 and bench.sampler_extra_steps(10.0, 0) == 0
    world = 2
    mp.spawn(_timed_loop_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "calls0.npy"), np.load(tmp_path / "calls1.npy")
    assert (a == b).all() and a[0] > 5
