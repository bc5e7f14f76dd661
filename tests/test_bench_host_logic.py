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
