"""Dry run of bench.main() on a CPU-only box with a fake torch.cuda and a fake engine module: catches NameError /
TypeError / logic slips in the N = 1 path of bench.py, including the extra steps of the clock sampler and the JSON
contract keys (no number it prints is meaningful).  Run as a subprocess by tests/test_bench_host_logic.py."""
import sys, types, ctypes, json, io, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

# ---- fake torch.cuda
class FakeStream: cuda_stream = 0
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self, stream=None): pass
    def elapsed_time(self, other): return 12.5 + 7.0 * RANK  # ranks disagree until the all-reduce
torch.cuda.set_device = lambda d: None
torch.cuda.current_stream = lambda: FakeStream()
torch.cuda.Event = FakeEvent
torch.cuda.synchronize = lambda: None
torch.cuda.empty_cache = lambda: None
_real_device = torch.device
def fake_device(kind, idx=None): return _real_device("cpu")
torch.device = fake_device
for name in ("empty", "zeros"):
    real = getattr(torch, name)
    def wrap(*a, _real=real, **k):
        k.pop("pin_memory", None)
        return _real(*a, **k)
    setattr(torch, name, wrap)
_real_tensor = torch.tensor
torch.tensor = lambda *a, **k: _real_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"} )

if WORLD > 1:  # N > 1 control flow on the CPU: gloo stands in for NCCL
    import torch.distributed as dist
    _real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: _real_init("gloo", rank=RANK, world_size=WORLD)

# ---- fake engine module
class FakeLib:
    def __getattr__(self, name):
        if name == "msm_last_error": return lambda h: b"fake"
        if name == "msm_field_impl": return lambda h: b"fake/field"
        if name == "msm_bases_table_window": return lambda b: 22
        if name == "msm_to_affine":
            def f(h, jac, count, mont, xy, inf): return 0
            return f
        return lambda *a, **k: 0
class FakeWs:
    def __init__(self, curve, devices=None): self.handle = ctypes.c_void_p(1); self.n = 0
    def timings(self):
        self.n += 13
        return {"sort_ms": 2.7, "accumulate_ms": 30.5, "reduce_ms": 1.5, "total_ms": 34.8, "window_bits": 22, "num_windows": 12,
                "kernel_launches": self.n, "scatter_passes": 0, "sub_batches": 1, "h2d_ms": 0.0, "num_entries": 1}
fake = types.ModuleType("ec_gpu_b200")
fake.load_library = lambda: FakeLib()
fake.fq_bytes = lambda c: 32 if c == 0 else 48
fake.Workspace = FakeWs
def shard_range(n, parts, i):
    c = (n + parts - 1) // parts
    return min(i * c, n), min((i + 1) * c, n)
fake.shard_range = shard_range
class FakeKern:
    @classmethod
    def create(cls, devs, curve): return cls()
    def multiexp(self, w, pts, sc, skip):
        from oracle import oracle as O
        return O.multiexp_cpu(0, pts, sc)
fake.MultiexpKernel = FakeKern
fake.Worker = lambda: None
sys.modules["ec_gpu_b200"] = fake

import bench
bench.ClockSampler.start = lambda self: setattr(self, "proc", None)
RUNS = ([["bench.py", "--gpus", str(WORLD), "--log-n", "12", "--steps", "2", "--warmup", "1"]] if WORLD > 1 else
        [["bench.py", "--log-n", "12", "--steps", "2", "--warmup", "1", "--cpu-log-sample", "12"],
         ["bench.py", "--workload", "batched", "--steps", "2", "--warmup", "1", "--no-cpu-baseline"]])
for argv in RUNS:
    sys.argv = argv
    r, w = os.pipe()
    saved = os.dup(1)
    os.dup2(w, 1)
    try:
        bench.main()
    finally:
        os.dup2(saved, 1)
        os.close(w)
    if RANK != 0:  # other ranks print nothing (and bench keeps a dup of the write end open: a read would block)
        os.close(r)
        continue
    out = os.read(r, 1 << 20).decode()
    line = [l for l in out.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print("OK", argv[1:3], sorted(d.keys()))
    assert {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"} <= set(d.keys())
    print("  gpu_launches", d["gpu_launches"], "roofline.frac %.3f" % d["roofline"]["frac"], "cpu_baseline" in d, d["clocks"])
