#!/usr/bin/env python3
"""Measurement of the SURVEY.md section 8f rows beside the headline bench: the batched / AMT call shapes,
the EC-FFT and the scalar-field FFT -- device time (CUDA events inside the engine), end-to-end time
through the reference-facing call with host buffers, the oracle's CPU restatement on a bounded sample
of the same workload, and the fraction of the IMAD roofline.  One JSON line per workload.

  python tests/perf/bench_next_rows.py [batched] [amt] [ecfft] [fft]

Lives under tests/ because it uses the oracle (as the checker and as the CPU baseline), which only tests/,
__graft_entry__.smoke() and bench.py may do.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ec_gpu_b200 as m  # noqa: E402
from oracle import oracle as O  # noqa: E402  (CPU baseline legs only)

IMAD_PEAK = 148 * 64 * 1.965e9
SEED = 0x0BADC0DE
R_FR = {0: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
        1: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001}
GEN = {0: 5, 1: 7}


def best_of(fn, reps=4):
    best = None
    for it in range(reps):
        t0 = time.perf_counter()
        dev_ms = fn()
        wall = (time.perf_counter() - t0) * 1e3
        if it and (best is None or wall < best[0]):
            best = (wall, dev_ms)
    return best


def chunked(lib, name, lines, log_L, chunks, canon_c):
    """multiple_multiexp shapes of ag-cuda-ec/benches/{multiexp,amt}.rs, BN254."""
    curve, fq = 0, 32
    ws = m.Workspace(curve)
    h = ws.handle
    L, n = 1 << log_L, lines << log_L
    dp, ds = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.msm_device_alloc(h, n * 2 * fq, ctypes.byref(dp)) == 0
    assert lib.msm_device_alloc(h, L * 32, ctypes.byref(ds)) == 0
    assert lib.msm_synth_points_device(h, SEED, 0, n, dp) == 0
    assert lib.msm_synth_scalars_device(h, SEED, 0, L, ds) == 0
    bh = ctypes.c_void_p()
    assert lib.msm_bases_from_device(h, dp, n, ctypes.byref(bh)) == 0
    assert lib.msm_bases_precompute_chunked(h, bh, L // chunks) == 0
    sc = np.zeros((L, 32), dtype=np.uint8)
    assert lib.msm_memcpy_d2h(h, sc.ctypes.data, ds, sc.nbytes) == 0
    assert lib.msm_host_register(sc.ctypes.data, sc.nbytes) == 0
    out = np.zeros((lines * chunks, 3 * fq), dtype=np.uint8)

    def call():
        rc = lib.msm_multiple_multiexp(h, bh, sc.ctypes.data, L, chunks, 8, 1, out.ctypes.data)
        assert rc == 0, lib.msm_last_error(h)
        return ws.timings()["total_ms"]

    wall, dev = best_of(call)
    t = ws.timings()
    # CPU: the oracle's multiple_multiexp on the first `sample` chunks of line 0
    sample = 32
    cl = L // chunks
    pts = np.zeros((sample * cl, 2 * fq), dtype=np.uint8)
    assert lib.msm_memcpy_d2h(h, pts.ctypes.data, dp, pts.nbytes) == 0
    t0 = time.perf_counter()
    ref = O.multiple_multiexp(curve, pts, sc[: sample * cl], sample)
    cpu_s = time.perf_counter() - t0
    same = bool((O.to_affine(curve, ref)[0] == O.to_affine(curve, out[:sample])[0]).all())
    W = (254 + 1 + canon_c - 1) // canon_c
    macs = W * 10 * 136
    lib.msm_host_unregister(sc.ctypes.data)
    return {"workload": name, "curve": "bn254", "lines": lines, "log_L": log_L, "num_chunks": chunks,
            "points_per_call": n, "window_bits": t["window_bits"], "num_windows": t["num_windows"],
            "device_ms": round(dev, 3), "value": n / (dev * 1e-3), "unit": "points/s",
            "e2e": {"ms": round(wall, 3), "value": n / (wall * 1e-3), "h2d_bytes": L * 32, "d2h_bytes": out.nbytes},
            "roofline": {"bound": "imad", "algorithmic_macs_per_point": macs, "canonical_window": canon_c,
                         "frac": n / (dev * 1e-3) * macs / IMAD_PEAK},
            "cpu_baseline": {"value": sample * cl / cpu_s, "unit": "points/s", "cores": O.ncores(), "kind": "port",
                             "sample": "oracle multiple_multiexp on the first %d chunks of line 0 (%.2f s)" % (sample, cpu_s)},
            "matches_oracle_on_sample": same}


def omegas_for(curve, n, single=False):
    r = R_FR[curve]
    omega = pow(GEN[curve], (r - 1) // n, r)
    if single:
        return np.frombuffer((omega * (1 << 256) % r).to_bytes(32, "little"), dtype=np.uint8).copy()
    out = np.zeros((32, 32), dtype=np.uint8)
    for i in range(32):
        out[i] = np.frombuffer((pow(omega, 1 << i, r) * (1 << 256) % r).to_bytes(32, "little"), dtype=np.uint8)
    return out


def ecfft(lib, log_n, cpu_log_n):
    curve, fq = 0, 32
    ws = m.Workspace(curve)
    n = 1 << log_n
    jac = np.zeros((n, 3 * fq), dtype=np.uint8)
    jac[:, :2 * fq] = O.gen_points(curve, SEED, n)
    jac[:, 2 * fq:] = O.constant(curve, 1)
    om = omegas_for(curve, n)
    work = jac.copy()

    def call():
        work[:] = jac
        m.radix_ec_fft(ws, work, om)
        return ws.timings()["total_ms"]

    wall, dev = best_of(call, 3)
    nc = 1 << cpu_log_n
    t0 = time.perf_counter()
    ref = O.ec_fft(curve, jac[:nc], omegas_for(curve, nc)[0])
    cpu_s = time.perf_counter() - t0
    chk = jac[:nc].copy()
    m.radix_ec_fft(ws, chk, omegas_for(curve, nc))
    same = bool((O.to_affine(curve, ref)[0] == O.to_affine(curve, chk)[0]).all())
    bf = (n // 2) * log_n
    products = 7 * 14 + 128 * 9 + 66 * 15 + 2 * 14  # GLV: table, 128 doublings, <= 66 additions (+ beta X), two butterfly additions
    return {"workload": "EC-FFT over G1 (radix_ec_fft), 2^%d points" % log_n, "curve": "bn254", "log_n": log_n,
            "device_ms": round(dev, 3), "value": bf / (dev * 1e-3), "unit": "butterflies/s",
            "e2e": {"ms": round(wall, 3), "value": bf / (wall * 1e-3), "h2d_bytes": jac.nbytes, "d2h_bytes": jac.nbytes},
            "roofline": {"bound": "imad", "algorithmic_macs_per_butterfly": products * 136,
                         "frac": bf / (dev * 1e-3) * products * 136 / IMAD_PEAK,
                         "note": "below 2^14 points the transform is latency-bound (log_n dependent scalar multiplications)"},
            "cpu_baseline": {"value": (nc // 2) * cpu_log_n / cpu_s, "unit": "butterflies/s", "cores": 1, "kind": "port",
                             "sample": "oracle serial_ec_fft on 2^%d of the points (%.2f s)" % (cpu_log_n, cpu_s)},
            "matches_oracle_on_sample": same}


def fft(lib, log_n, cpu_log_n):
    curve = 0
    k = m.FftKernel.create([0], curve)
    ws = k.workspace
    n = 1 << log_n
    a = O.gen_scalars(curve, SEED, n)  # canonical values < r are valid Montgomery residues
    om = omegas_for(curve, n, single=True)
    assert lib.msm_host_register(a.ctypes.data, a.nbytes) == 0
    work = a.copy()
    assert lib.msm_host_register(work.ctypes.data, work.nbytes) == 0

    def call():
        work[:] = a
        t0 = time.perf_counter()
        k.radix_fft(work, om, log_n)
        call.inner = (time.perf_counter() - t0) * 1e3
        return ws.timings()["total_ms"]

    best = None
    for it in range(3):
        dev = call()
        if it and (best is None or call.inner < best[0]):
            best = (call.inner, dev)
    wall, dev = best
    nc = 1 << cpu_log_n
    t0 = time.perf_counter()
    ref = O.fr_fft(curve, a[:nc], omegas_for(curve, nc, single=True))
    cpu_s = time.perf_counter() - t0
    chk = a[:nc].copy()
    k.radix_fft(chk, omegas_for(curve, nc, single=True), cpu_log_n)
    bf = (n // 2) * log_n
    lib.msm_host_unregister(a.ctypes.data)
    lib.msm_host_unregister(work.ctypes.data)
    return {"workload": "scalar-field FFT (radix_fft), 2^%d elements" % log_n, "curve": "bn254 Fr", "log_n": log_n,
            "device_ms": round(dev, 3), "value": n / (dev * 1e-3), "unit": "elements/s",
            "e2e": {"ms": round(wall, 3), "value": n / (wall * 1e-3), "h2d_bytes": a.nbytes, "d2h_bytes": a.nbytes},
            "roofline": {"bound": "imad", "algorithmic_macs_per_butterfly": 136, "frac": bf / (dev * 1e-3) * 136 / IMAD_PEAK,
                         "hbm_frac_at_4_passes": (n * 64 * 4) / (dev * 1e-3) / 6546.6e9},
            "cpu_baseline": {"value": nc / cpu_s, "unit": "elements/s", "cores": 1, "kind": "port",
                             "sample": "oracle serial_fft on 2^%d elements (%.2f s)" % (cpu_log_n, cpu_s)},
            "matches_oracle_on_sample": bool((ref == chk).all())}


def main():
    lib = m.load_library()
    O.build()
    which = set(sys.argv[1:]) or {"batched", "amt", "ecfft", "fft"}
    rows = []
    if "batched" in which:
        rows.append(chunked(lib, "1024 x 2^12 per-segment commitments (ag-cuda-ec/benches/multiexp.rs:19-22,56)", 1, 22, 1024, 8))
    if "amt" in which:
        rows.append(chunked(lib, "AMT: 10 lines x 2^21, 2048 chunks (ag-cuda-ec/benches/amt.rs:18-55)", 10, 21, 2048, 8))
    if "ecfft" in which:
        rows.append(ecfft(lib, 11, 7))
        rows.append(ecfft(lib, 16, 7))
    if "fft" in which:
        rows.append(fft(lib, 20, 18))
        rows.append(fft(lib, 24, 18))
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
