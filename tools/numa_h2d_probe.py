#!/usr/bin/env python3
"""Where does the end-to-end time of an N-GPU step go?  (development aid; bench.py is the contract)

One process per GPU under torch.distributed.run.  Every rank holds a shard of 2^log_n / N BN254 points with its window
table and measures, with all ranks working at the same time (barrier before every sample):

  * the host -> device copy of its scalar shard alone (GB/s), from a pinned buffer allocated (a) wherever the process
    happened to run and (b) after binding the process to the CPUs of the GPU's NUMA node (sysfs `local_cpulist`);
  * the same copy with the other ranks idle (rank by rank);
  * msm_multiple_multiexp from each of the two buffers, for upload pipeline depths 1, 2, 3, 4.

  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29533 tools/numa_h2d_probe.py 24 > gpurun_out/numa_probe.jsonl
"""
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ec_gpu_b200 as m  # noqa: E402
from bench import gpu_numa_cpus  # noqa: E402

SEED = 0x0BADC0DE


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = (1 << log_n) // world
    lib = m.load_library()
    ws = m.Workspace(0, devices=[local])
    h = ws.handle
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    d_pts = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    d_sc = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    assert lib.msm_synth_points_device(h, SEED, rank * n, n, p(d_pts)) == 0
    assert lib.msm_synth_scalars_device(h, SEED, rank * n, n, p(d_sc)) == 0
    bh = ctypes.c_void_p()
    assert lib.msm_bases_from_device(h, p(d_pts), n, ctypes.byref(bh)) == 0
    del d_pts
    assert lib.msm_bases_set_table_policy(h, bh, 2) == 0
    h_out = torch.zeros(96, dtype=torch.uint8, pin_memory=True)
    d_tmp = torch.empty(n * 32, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def say(**kw):
        kw.update(rank=rank, world=world, log_n=log_n)
        print(json.dumps(kw), flush=True)

    cpus0 = sorted(os.sched_getaffinity(0))
    info = gpu_numa_cpus(local)
    say(what="topology", affinity_before=len(cpus0), gpu_numa=info)
    bufs = {}
    bufs["default"] = torch.empty(n * 32 + 4096, dtype=torch.uint8, pin_memory=True)[: n * 32]
    bufs["default"].copy_(d_sc)
    if info and info.get("cpus"):
        try:
            os.sched_setaffinity(0, info["cpus"])
            bufs["numa_local"] = torch.empty(n * 32 + 8192, dtype=torch.uint8, pin_memory=True)[: n * 32]
            bufs["numa_local"].copy_(d_sc)
        except OSError as e:
            say(what="setaffinity failed", error=str(e))
    torch.cuda.synchronize()

    def copy_gbs(buf, reps=8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            d_tmp.copy_(buf, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return buf.numel() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    for name, buf in bufs.items():
        copy_gbs(buf, 2)
        barrier()
        say(what="h2d all ranks at once", buffer=name, gbs=round(copy_gbs(buf), 2), mbytes=buf.numel() >> 20)
        for r in range(world):
            barrier()
            if r == rank:
                say(what="h2d alone", buffer=name, gbs=round(copy_gbs(buf), 2))
        barrier()

    # the call: device scalars first (the floor), then host scalars from each buffer and pipeline depth
    def call(src_ptr, device):
        if device:
            d_o = torch.empty(96, dtype=torch.uint8, device=dev)
            rc = lib.msm_multiple_multiexp_device(h, bh, src_ptr, n, 1, p(d_o))
        else:
            rc = lib.msm_multiple_multiexp(h, bh, src_ptr, n, 1, 8, 1, p(h_out))
        assert rc == 0, lib.msm_last_error(h)
        torch.cuda.synchronize()

    def timed_call(src_ptr, device, reps=6):
        call(src_ptr, device)
        call(src_ptr, device)
        ts = []
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            call(src_ptr, device)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        t = ws.timings()
        return {"ms_median": round(ts[len(ts) // 2], 3), "ms_min": round(ts[0], 3), "device_total_ms": round(t["total_ms"], 3),
                "h2d_ms": round(t["h2d_ms"], 3), "sub_batches": t["sub_batches"], "window_bits": t["window_bits"]}

    say(what="call, device scalars", **timed_call(p(d_sc), True))
    for name, buf in bufs.items():
        for depth in (0, 1, 2, 3, 4):
            if depth:
                os.environ["MSM_B200_PIPELINE"] = str(depth)
            else:
                os.environ.pop("MSM_B200_PIPELINE", None)
            say(what="call, host scalars", buffer=name, pipeline=depth or "default", **timed_call(p(buf), False))
    os.environ.pop("MSM_B200_PIPELINE", None)
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
