#!/usr/bin/env python3
"""bench.py -- BN254 G1 MSM throughput (BASELINE.json: "BN254 G1 MSM points/sec (2^24, 1/2/4/8 B200)
+ % IMAD roofline vs host CPU").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 24] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one complete MSM of 2^log_n synthetic points (seed 0x0badc0de).
With N GPUs the point/scalar arrays are split into N contiguous shards (the partition of
MultiexpKernel::parallel_multiexp, ec-gpu-proxy/src/multiexp.rs:329-337), each rank computes one
partial point, the N partials (96 bytes each) are all-gathered over NCCL/NVLink and summed on the
device.  Total work is fixed => "scaling": "strong".

  value     whole-job points/s with bases AND scalars resident in HBM (msm_multiple_multiexp_device)
  e2e       the same through the reference-facing call with HOST scalars (pinned) and a host result:
            msm_multiple_multiexp == ag_cuda_ec::multiple_multiexp (bases resident, as in that API)
  roofline  the dominant kernel (k_accumulate, bucket accumulation) against the int32-multiply
            (IMAD) pipe: algorithmic MACs per launch = n * 21760 (16 windows x 10 field products x
            136 MACs) / its CUDA-event time on the launching stream
  cpu_baseline  the oracle's restatement of the reference's multiexp_cpu on the host cores, on a
            bounded sample (the only place this file touches oracle/ besides --impl reference)

Timing: CUDA events on the stream every kernel is launched on (the engine is switched onto torch's
current stream), barrier + synchronize on both sides, max over ranks.  Inputs per step (1 GiB of
bases, 0.5 GiB of scalars, >= 0.8 GiB of sorted digits) exceed the 126 MB L2 many times over.
Clocks: nvidia-smi sampled every 100 ms from before the warm-up; when the timed region is shorter than
that, the same step keeps running for ~0.3 s after it -- the SAME number of extra steps on every rank
(sampler_extra_steps: a function of the all-reduced time), because a step holds a collective.

  --workload batched   BASELINE.json configs[4]: 1024 BN254 MSMs of 2^12 points (chunked window table),
                       tasks split over the ranks, results all-gathered, nothing to add up
  --curve 1 --log-n 22 BASELINE.json configs[3]: BLS12-381 G1
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x0BADC0DE
MACS_PER_POINT = {0: 16 * 10 * 136, 1: 16 * 10 * 300}  # SURVEY.md section 8d / BASELINE.md section 3
IMAD_PEAK_NOMINAL = 148 * 64 * 1.965e9  # MAC/s; tools/imad_peak.cu measured 1.852e13 (99.5 %) on this pool
METRIC = {0: "BN254 G1 MSM points/sec", 1: "BLS12-381 G1 MSM points/sec"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def stop(self, t0=None, t1=None):
        """Rows stamped inside [t0, t1] (the timed region); all rows when no window is given."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        self.thread.join(timeout=2)
        window = "timed region"
        picked = [r[1:] for r in self.rows if t0 is None or (t0 <= r[0] <= t1 + 0.02)]
        if len(picked) < 2:
            # a region shorter than the sampling period (N = 8: 5 steps take 28 ms): use every sample
            # taken under the same load -- warm-up, timed region and the burst that follows it
            picked = [r[1:] for r in self.rows]
            window = "warm-up + timed region + post-region burst of the same step (region shorter than the sampling period)"
        self.rows = picked
        self.window = window
        def num(s):
            try:
                return float(s)
            except ValueError:
                return None

        rows = [r for r in self.rows if len(r) >= 7 and num(r[0]) is not None]
        sm = [num(r[0]) for r in rows]
        mx = [num(r[1]) for r in rows if num(r[1]) is not None]
        pw = [num(r[2]) or 0.0 for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        busy = [c for c, p in zip(sm, pw) if p > 250] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons, "window": window}


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the sysfs cpulist format)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(device_index):
    """The NUMA node of a GPU and the CPUs next to it, from sysfs (None when the box does not say).  A pinned host
    buffer allocated by a thread running on those CPUs lands on the memory of the socket the GPU's PCIe link hangs
    off, so N ranks uploading at once do not all pull through one socket's memory and the inter-socket link."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device_index)
        addr = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + addr
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        return {"pci": addr, "node": node, "cpus": cpus}
    except Exception:  # noqa: BLE001 -- sysfs not exposed, attribute missing: no binding, nothing else changes
        return None


def sampler_extra_steps(ms_total, steps):
    """Untimed steps to append after a timed region of ms_total milliseconds (max over ranks) so that the
    clock sampler sees at least ~0.3 s of the same load.  A pure function of values that are identical on
    every rank (tests/test_sharding_gloo.py pins that property): the step contains a collective."""
    if ms_total >= 250.0 or steps <= 0:
        return 0
    per_step = max(ms_total / steps, 0.05)
    return min(int(-(-300.0 // per_step)), 2000)


def cpu_baseline(curve, log_sample, threads=None):
    """The oracle's multiexp_cpu (restatement of ec-gpu-proxy/src/multiexp_cpu.rs:244-367) on the
    host cores: a reported baseline, not the target."""
    from oracle import oracle as O

    O.build()
    n = 1 << log_sample
    threads = threads or O.ncores()
    pts = O.gen_points(curve, SEED, n)
    sc = O.gen_scalars(curve, SEED, n)
    t0 = time.perf_counter()
    O.multiexp_cpu(curve, pts, sc, nthreads=threads)
    dt = time.perf_counter() - t0
    c = O.window_for(n)
    bits = 254 if curve == 0 else 255
    windows = (bits + c - 1) // c
    return {"value": n / dt, "unit": "points/s", "cores": min(threads, windows), "kind": "port",
            "host_threads": threads, "seconds": round(dt, 3),
            "sample": "one multiexp_cpu call on the first 2^%d points of the same synthetic workload "
                      "(c = %d, %d windows, one thread per window as the reference's rayon loop)" % (log_sample, c, windows)}


def run_reference(args, rank):
    """--impl reference: the reference's CPU multiexp (oracle port; the Rust original cannot be built
    here: no Rust toolchain, arkworks not vendored) with all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O

    O.build()
    curve = args.curve
    log_s = min(args.log_n, args.ref_log_sample)
    n = 1 << log_s
    pts = O.gen_points(curve, SEED, n)
    sc = O.gen_scalars(curve, SEED, n)
    threads = O.ncores()
    for _ in range(args.warmup):
        O.multiexp_cpu(curve, pts[: max(n // 8, 1)], sc[: max(n // 8, 1)], nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.multiexp_cpu(curve, pts, sc, nthreads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    c = O.window_for(n)
    bits = 254 if curve == 0 else 255
    windows = (bits + c - 1) // c
    name = "BN254 G1" if curve == 0 else "BLS12-381 G1"
    sample = ("each step = one multiexp_cpu call on the first 2^%d of the 2^%d synthetic points "
              "(c = %d, %d windows = usable threads)" % (log_s, args.log_n, c, windows))
    print(json.dumps({
        "impl": "reference", "metric": METRIC[curve], "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "%s MSM 2^%d (reference CPU multiexp, bounded sample per step)" % (name, args.log_n),
                   "log_n": args.log_n, "sample_log_n": log_s, "seed": SEED},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": min(threads, windows),
                         "host_threads": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(curve, log_n, n_local):
    """dram bytes per k_accumulate launch from the committed ncu --set full capture of the same
    workload (profiles/ncu_traffic.json); null when no capture matches this configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            for e in json.load(f)["captures"]:
                if e["curve"] == curve and e["log_n"] == log_n and e["points_per_gpu"] == n_local:
                    return e["k_accumulate_dram_bytes_per_launch"], e["source"]
    except (OSError, ValueError, KeyError):
        pass
    return None, None


def to_affine_bytes(lib, h, jac, fq):
    import numpy as np

    count = jac.size // (3 * fq)
    xy = np.zeros(count * 2 * fq, dtype=np.uint8)
    inf = np.zeros(count, dtype=np.uint8)
    assert lib.msm_to_affine(h, jac.ctypes.data, count, 0, xy.ctypes.data, inf.ctypes.data) == 0
    return np.concatenate([xy, inf])


def golden_check(curve, log_n, batched, affine_bytes, fq):
    """The final result (after the cross-GPU sum, at every N) against tests/golden/fullsize.json -- the CPU
    oracle's canonical affine result on the same seeded inputs, committed; no oracle code runs here."""
    import numpy as np

    try:
        with open(os.path.join(ROOT, "tests", "golden", "fullsize.json")) as f:
            g = json.load(f)
    except (OSError, ValueError):
        return None, None
    if batched:
        key, rows = "bn254_batched_1024x4096", None
        rows = g.get(key, {}).get("results")
    else:
        key = ("bn254_2p%d" if curve == 0 else "bls12_381_2p%d") % log_n
        rows = [g[key]["result"]] if key in g else None
    if rows is None:
        return None, None
    count = len(rows)
    want_xy = b"".join(bytes.fromhex(r["x"])[::-1] + bytes.fromhex(r["y"])[::-1] for r in rows)
    want_inf = bytes(r["inf"] for r in rows)
    want = np.frombuffer(want_xy + want_inf, dtype=np.uint8)
    return bool(affine_bytes.size == want.size and count * (2 * fq + 1) == want.size and (affine_bytes == want).all()), key


def spot_check(m, curve, device):
    """Untimed: a 2^14-point MSM of the same synthetic stream through the public API equals the
    oracle's result bit-exactly (canonical affine)."""
    from oracle import oracle as O

    n = 1 << 14
    pts, sc = O.gen_points(curve, SEED, n), O.gen_scalars(curve, SEED, n)
    kern = m.MultiexpKernel.create([device], curve)
    got = kern.multiexp(m.Worker(), pts, sc, 0)
    want = O.multiexp_cpu(curve, pts, sc)
    ga, gi = O.to_affine(curve, got)
    wa, wi = O.to_affine(curve, want)
    return bool((ga == wa).all() and (gi == wi).all())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--curve", type=int, default=0, help="0 = BN254 G1 (headline), 1 = BLS12-381 G1")
    ap.add_argument("--cpu-log-sample", type=int, default=23, help="cpu_baseline sample size (log2)")
    ap.add_argument("--ref-log-sample", type=int, default=24,
                    help="--impl reference sample per step (log2): 2^24 = the whole workload (c = 17, 15 windows <= host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="msm", choices=["msm", "batched"],
                    help="msm: one MSM of 2^log_n points (the headline, BASELINE.json configs[1-3]); batched: 1024 "
                         "independent BN254 MSMs of 2^12 points (configs[4], ag-cuda-ec/benches/multiexp.rs:19-22,56), "
                         "the tasks split over the ranks, no reduction")
    ap.add_argument("--no-table", action="store_true",
                    help="table policy off for the headline too (the no_table leg always runs with it off)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if rank == 0:
            print("bench.py: --gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d"
                  % (args.gpus, world, args.gpus), file=sys.stderr)
        sys.exit(2)

    # stdout carries exactly one JSON line: anything libraries print there while we run (NCCL's
    # version banner, for one) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import ec_gpu_b200 as m

    lib = m.load_library()  # ImportError if the CUDA library is missing: there is no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime

        # a collective that cannot complete should fail the run in minutes, not hold 8 GPUs for NCCL's default 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    curve = args.curve
    fq = m.fq_bytes(curve)
    ws = m.Workspace(curve, devices=[local_rank])
    h = ws.handle
    stream = torch.cuda.current_stream()
    assert lib.msm_set_stream(h, ctypes.c_void_p(stream.cuda_stream)) == 0

    batched = args.workload == "batched"
    if batched:
        args.log_n, chunk_len = 22, 4096
        assert curve == 0 and (1024 % world) == 0, "batched workload: BN254, 1024 tasks split evenly over the ranks"
    n_total = 1 << args.log_n
    start, end = m.shard_range(n_total, world, rank)
    n_local = end - start
    chunks_local = n_local // chunk_len if batched else 1  # tasks of this rank
    macs_per_point = 32 * 10 * 136 if batched else MACS_PER_POINT[curve]  # canonical c = 8 for 4096-point MSMs
    d_pts = torch.empty(n_local * 2 * fq, dtype=torch.uint8, device=dev)
    d_sc = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(chunks_local * 3 * fq, dtype=torch.uint8, device=dev)
    d_gather = torch.zeros(world * chunks_local * 3 * fq, dtype=torch.uint8, device=dev)
    d_final = torch.zeros((world if batched else 1) * chunks_local * 3 * fq, dtype=torch.uint8, device=dev)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    assert lib.msm_synth_points_device(h, SEED, start, n_local, ptr(d_pts)) == 0
    assert lib.msm_synth_scalars_device(h, SEED, start, n_local, ptr(d_sc)) == 0
    # The drop-in call sequence and nothing else: upload (here from device memory, the points were generated
    # there) -> multiple_multiexp.  No engine-only call in between: the window table appears by policy on the
    # second call of the shape, i.e. during the warm-up (include/msm_b200.h, "Window tables by policy").
    bases, bases_plain = ctypes.c_void_p(), ctypes.c_void_p()
    t_setup = time.perf_counter()
    rc = lib.msm_bases_from_device(h, ptr(d_pts), n_local, ctypes.byref(bases))
    assert rc == 0, lib.msm_last_error(h)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    if args.no_table:
        assert lib.msm_bases_set_table_policy(h, bases, 0) == 0
    # second resident copy with the table policy off: the no_table leg (what a one-off call gets)
    rc = lib.msm_bases_from_device(h, ptr(d_pts), n_local, ctypes.byref(bases_plain))
    assert rc == 0, lib.msm_last_error(h)
    assert lib.msm_bases_set_table_policy(h, bases_plain, 0) == 0
    del d_pts
    torch.cuda.empty_cache()
    h_sc = torch.empty(n_local * 32, dtype=torch.uint8, pin_memory=True)
    h_sc.copy_(d_sc)
    h_out = torch.zeros(d_final.numel(), dtype=torch.uint8, pin_memory=True)
    h_part = torch.zeros(chunks_local * 3 * fq, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def combine():
        if world > 1:
            dist.all_gather_into_tensor(d_gather, d_out)
            if batched:
                d_final.copy_(d_gather)  # rank-major = task order: nothing to add up
            elif rank == 0:
                assert lib.msm_sum_points_device(h, ptr(d_gather), world, ptr(d_final)) == 0
        else:
            d_final.copy_(d_out)

    acc_ms = []

    def step_device():
        rc = lib.msm_multiple_multiexp_device(h, bases, ptr(d_sc), n_local, chunks_local, ptr(d_out))
        assert rc == 0, lib.msm_last_error(h)
        acc_ms.append(ws.timings())
        combine()

    def step_plain():
        rc = lib.msm_multiple_multiexp_device(h, bases_plain, ptr(d_sc), n_local, chunks_local, ptr(d_out))
        assert rc == 0, lib.msm_last_error(h)
        combine()

    def step_e2e():
        dst = h_part if world > 1 else h_out
        rc = lib.msm_multiple_multiexp(h, bases, ctypes.c_void_p(h_sc.data_ptr()), n_local, chunks_local, 8, 1,
                                       ctypes.c_void_p(dst.data_ptr()))
        assert rc == 0, lib.msm_last_error(h)
        if world > 1:
            d_out.copy_(h_part, non_blocking=True)
            combine()
            if rank == 0:
                h_out.copy_(d_final)
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        clocks = None
        if sampler is not None:  # every rank passes one (rank-uniform branch: it holds collectives)
            # Keep the same load up until nvidia-smi (100 ms period) has seen it.  The step holds a collective
            # when world > 1, so every rank must run the SAME number of extra steps: the count comes from the
            # all-reduced time, never from a local clock.
            for _ in range(sampler_extra_steps(ms, steps)):
                fn()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            clocks = sampler.stop(t0, t1)
        return ms, clocks

    launches0 = None
    # warm-up happens inside timed(); launches are counted over the timed steps only
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs ~0.1 s to deliver its first row: started before the warm-up
    for _ in range(args.warmup):
        step_device()
    launches0 = ws.timings()["kernel_launches"]
    del acc_ms[:]
    ms_dev, clocks = timed(step_device, args.steps, 0, sampler=sampler)
    # the sampler may have kept the load up after the region: only the K timed steps count
    t_last = acc_ms[args.steps - 1]
    acc_timed = [t["accumulate_ms"] for t in acc_ms[:args.steps]]
    launches = t_last["kernel_launches"] - launches0
    result_dev = d_final.cpu().numpy().copy() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, min(args.warmup, 2)))
    e2e_sub_batches = int(ws.timings()["sub_batches"])
    result_e2e = h_out.numpy().copy() if rank == 0 else None
    ms_plain, _ = timed(step_plain, args.steps, 2)
    result_plain = d_final.cpu().numpy().copy() if rank == 0 else None

    if rank == 0:
        value = n_total * args.steps / (ms_dev * 1e-3)
        e2e = n_total * args.steps / (ms_e2e * 1e-3)
        # dominant kernel: bucket accumulation of this rank's shard
        acc_avg_ms = sum(acc_timed) / len(acc_timed)
        macs = n_local * macs_per_point
        achieved = macs / (acc_avg_ms * 1e-3)
        roofline = {"bound": "imad", "kernel": "k_accumulate", "achieved": achieved / 1e12,
                    "peak": IMAD_PEAK_NOMINAL / 1e12, "unit": "TMAC/s", "frac": achieved / IMAD_PEAK_NOMINAL,
                    "traffic": None, "avg_kernel_ms": acc_avg_ms, "kernel_ms_steps": [round(x, 3) for x in acc_timed],
                    "algorithmic_macs_per_launch": macs,
                    "peak_source": "148 SMs x 64 int32-multiply lanes/clk x 1.965 GHz; tools/imad_peak.cu measured "
                                   "1.852e13 MAC/s (99.5 % of it) on this pool; MEASURED_PEAKS.json has no integer figure",
                    "whole_step_frac": value / world * macs_per_point / IMAD_PEAK_NOMINAL}
        traffic, traffic_src = ncu_traffic(curve, args.log_n, n_local)
        roofline["traffic"] = traffic
        roofline["traffic_source"] = traffic_src
        roofline["algorithmic_gather_bytes_per_launch"] = n_local * t_last["num_windows"] * 2 * fq
        # sort phase (digit decomposition + histogram + scatter) against HBM
        peaks, peak_src = measured_peaks()
        binned = t_last["scatter_passes"] == 0
        if binned:
            # binned sort: scalars read by k_bin_count and once by k_partition (digits stay in registers between its
            # two passes); 8-byte (bucket, entry) pairs written once and read by k_bin_hist (bucket ids: 4 B) and
            # k_bin_place (8 B); sorted entries written once
            sort_bytes = n_local * 32 * 2 + n_local * t_last["num_windows"] * (8 + 4 + 8 + 4)
            kernels = "k_bin_count + k_partition + k_bin_hist + scan + k_bin_place"
            note = ("32 B/scalar read by 2 passes + per digit: 8 B written and 12 B read of the partitioned "
                    "(bucket, entry) pairs + 4 B of the sorted entry; all per-digit atomics are in shared memory, "
                    "the phase is bound by shared-memory atomics and store transactions, not by HBM bandwidth")
        else:
            sort_bytes = n_local * 32 * (1 + t_last["scatter_passes"]) + n_local * t_last["num_windows"] * 4
            kernels = "k_digits<count> + scan + k_digits<scatter> x passes"
            note = ("32 B/scalar read per pass + 4 B written per digit; the phase is bound by L2 atomics and "
                    "4-byte scattered writes, not by HBM bandwidth")
        sort_gbs = sort_bytes / (t_last["sort_ms"] * 1e-3) / 1e9 if t_last["sort_ms"] > 0 else None
        roofline_sort = {"bound": "hbm", "kernels": kernels,
                         "achieved": sort_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": (sort_gbs / peaks["hbm_gbs"]) if sort_gbs else None, "peak_source": peak_src,
                         "algorithmic_bytes": sort_bytes, "phase_ms": t_last["sort_ms"], "note": note}
        aff_dev = to_affine_bytes(lib, h, result_dev, fq)
        same = bool((aff_dev == to_affine_bytes(lib, h, result_e2e, fq)).all() and
                    (aff_dev == to_affine_bytes(lib, h, result_plain, fq)).all())
        matches_golden, golden_key = golden_check(curve, args.log_n, batched, aff_dev, fq)
        name = "BN254 G1" if curve == 0 else "BLS12-381 G1"
        cfg = {(0, 24): "configs[2]", (0, 20): "configs[1]", (1, 22): "configs[3]"}.get((curve, args.log_n),
                                                                                        "a size outside configs")
        out = {
            "metric": ("BN254 G1 batched MSM points/sec" if batched else METRIC[curve]), "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": ("1024 x BN254 G1 MSMs of 2^12 points, %d tasks per GPU, chunked window table "
                                    "(BASELINE.json configs[4]; roofline at the canonical c = 8 count, 43 520 MAC/point)"
                                    % chunks_local) if batched else
                                   "%s MSM 2^%d, contiguous shards of 2^%d points per GPU (BASELINE.json %s)"
                                   % (name, args.log_n, n_local.bit_length() - 1, cfg),
                       "log_n": args.log_n, "points_per_gpu": n_local, "window_bits": t_last["window_bits"],
                       "num_windows": t_last["num_windows"], "field_impl": lib.msm_field_impl(h).decode(),
                       "seed": SEED,
                       "call_sequence": "msm_bases_from_device (= upload_multiexp_bases, source already on the device) -> "
                                        "msm_multiple_multiexp[_device] x (warmup + steps); window table built by the "
                                        "engine's lazy policy on the 2nd call of the shape, inside the warm-up",
                       "window_table": int(lib.msm_bases_table_window(bases)) != 0,
                       "table_window_bits": int(lib.msm_bases_table_window(bases)),
                       "resident_setup_s": round(setup_s, 3),
                       "l2": "per-step inputs (bases %d MiB + scalars %d MiB + sorted digits) exceed the 126 MB L2"
                             % (n_local * 2 * fq >> 20, n_local * 32 >> 20)},
            "e2e": {"value": e2e, "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": n_total * 32, "d2h_bytes_per_step": world * chunks_local * 3 * fq,
                    "upload_sub_batches": e2e_sub_batches,
                    "call": "msm_multiple_multiexp: host scalars (pinned) in, host point out; bases resident as in "
                            "ag_cuda_ec::multiple_multiexp"},
            "gpu_launches": int(launches + (args.steps if world > 1 else 0)),
            "roofline": roofline,
            "roofline_sort": roofline_sort,
            "phases_ms": {k: round(t_last[k], 3) for k in ("sort_ms", "accumulate_ms", "reduce_ms", "total_ms")},
            "clocks": clocks,
            "paths_agree": same,
            "result_matches_golden": matches_golden,
            "golden": golden_key,
            "no_table": {"ms_per_step": ms_plain / args.steps, "value": n_total * args.steps / (ms_plain * 1e-3),
                         "unit": "points/s",
                         "frac": n_total * args.steps / (ms_plain * 1e-3) / world * macs_per_point / IMAD_PEAK_NOMINAL,
                         "note": "same call on a resident copy with the table policy off (what the first call of a "
                                 "shape, or a one-off call, gets); frac = whole-step fraction of the IMAD roofline"},
        }
        if not args.no_cpu_baseline:
            if world == 1:  # the CPU baseline is reported at N = 1 only (it does not depend on N)
                out["cpu_baseline"] = cpu_baseline(curve, min(args.cpu_log_sample, args.log_n))
            out["spot_check_vs_oracle"] = spot_check(m, curve, local_rank)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
