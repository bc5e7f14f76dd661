"""bench.py's host logic without a GPU: the whole main() of the N = 1 path runs against a fake torch.cuda and a
fake engine in a subprocess (tests/perf/bench_dry_run.py), for the headline and the batched workload; the JSON
line must carry every key of the bench contract."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_bench_main_dry_run():
    r = subprocess.run([sys.executable, os.path.join(HERE, "perf", "bench_dry_run.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK [") == 2, r.stdout[-2000:]


def test_reference_arm_runs_on_cpu():
    """bench.py --impl reference: the CPU arm prints the contract line and exits 0 (bounded sample)."""
    import json

    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(HERE), "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--ref-log-sample", "12"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0


def test_bench_two_rank_control_flow_on_gloo():
    """The N > 1 control flow of bench.py (barriers, all-gather inside every step, all-reduced time, the clock
    sampler's extra steps) with two CPU ranks whose local timings disagree: must finish -- the version that took
    the extra-step count from a local clock dead-locked NCCL at N = 8."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "perf", "bench_dry_run.py")],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK [") == 1, r.stdout[-2000:]


def test_golden_check_accepts_the_golden_and_rejects_anything_else():
    """bench.golden_check: the bytes msm_to_affine returns for the golden point(s) compare equal, any flipped bit or a
    missing row does not, unknown sizes give None (no claim)."""
    import json

    import numpy as np

    sys.path.insert(0, os.path.dirname(HERE))
    import bench

    g = json.load(open(os.path.join(HERE, "golden", "fullsize.json")))

    def affine_bytes(rows):
        xy = b"".join(bytes.fromhex(r["x"])[::-1] + bytes.fromhex(r["y"])[::-1] for r in rows)
        return np.frombuffer(xy + bytes(r["inf"] for r in rows), dtype=np.uint8).copy()

    one = affine_bytes([g["bn254_2p24"]["result"]])
    assert bench.golden_check(0, 24, False, one, 32) == (True, "bn254_2p24")
    bad = one.copy()
    bad[5] ^= 1
    assert bench.golden_check(0, 24, False, bad, 32)[0] is False
    assert bench.golden_check(0, 24, False, one[:-1], 32)[0] is False
    assert bench.golden_check(0, 19, False, one, 32) == (None, None)
    bls = affine_bytes([g["bls12_381_2p22"]["result"]])
    assert bench.golden_check(1, 22, False, bls, 48) == (True, "bls12_381_2p22")
    rows = g["bn254_batched_1024x4096"]["results"]
    assert bench.golden_check(0, 22, True, affine_bytes(rows), 32) == (True, "bn254_batched_1024x4096")
    assert bench.golden_check(0, 22, True, affine_bytes(rows[::-1]), 32)[0] is False


def test_cpulist_and_numa_lookup_degrade_quietly():
    """bench.parse_cpulist reads sysfs' cpulist format; gpu_numa_cpus returns None where there is no GPU / no sysfs entry
    (it only ever informs where a pinned buffer is allocated)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import bench

    assert bench.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert bench.parse_cpulist("5") == [5] and bench.parse_cpulist("") == []
    assert bench.gpu_numa_cpus(0) is None  # no GPU in this container


def test_clock_sampler_summary_without_nvidia_smi(monkeypatch):
    """ClockSampler: rows inside the timed window are picked; a region shorter than the sampling period falls back to
    every sample under load; throttle reasons are reported by name; no nvidia-smi -> says so."""
    sys.path.insert(0, os.path.dirname(HERE))
    import bench

    class FakeProc:
        def terminate(self):
            pass

    class FakeThread:
        def join(self, timeout=None):
            pass

    def sampler(rows):
        s = bench.ClockSampler(0)
        s.proc, s.thread, s.rows = FakeProc(), FakeThread(), rows
        return s

    idle = ["345", "1965", "140.0", "Not Active", "Not Active", "Not Active", "Not Active"]
    busy = ["1965", "1965", "640.5", "Not Active", "Not Active", "Not Active", "Active"]
    rows = [[0.0] + idle, [1.0] + busy, [1.1] + busy, [1.2] + busy, [2.0] + idle]
    c = sampler([list(r) for r in rows]).stop(0.95, 1.25)
    assert c["window"] == "timed region" and c["samples"] == 3 and c["sm_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    c = sampler([list(r) for r in rows]).stop(1.04, 1.06)  # no sample inside: every sample, idle ones filtered by power
    assert c["window"].startswith("warm-up") and c["samples"] == 5 and c["sm_mhz"] == 1965.0 and c["power_w_max"] == 640.5
    s = bench.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]
