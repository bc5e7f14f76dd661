#!/usr/bin/env python3
"""Device-resident timing sweep through the C ABI (development aid; bench.py is the contract)."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ec_gpu_b200 as m  # noqa: E402


def main():
    lib = m.load_library()
    curve = int(os.environ.get("CURVE", "0"))
    sizes = [int(x) for x in (sys.argv[1:] or ["16", "20", "22", "24"])]
    windows = [int(x) for x in os.environ.get("WINDOWS", "0").split(",")]
    ws = m.Workspace(curve)
    h = ws.handle
    print("field impl:", lib.msm_field_impl(h).decode())
    fq = m.fq_bytes(curve)
    chunks = int(os.environ.get("CHUNKS", "1"))   # tasks per line (ag_cuda_ec::multiple_multiexp num_chunks)
    lines = int(os.environ.get("LINES", "1"))     # base lines sharing the scalar row (AMT shape)
    for lg in sizes:
        L = 1 << lg
        n = L * lines
        dp, ds, do = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        assert lib.msm_device_alloc(h, n * 2 * fq, ctypes.byref(dp)) == 0
        assert lib.msm_device_alloc(h, L * 32, ctypes.byref(ds)) == 0
        assert lib.msm_device_alloc(h, 3 * fq * chunks * lines, ctypes.byref(do)) == 0
        t0 = time.time()
        assert lib.msm_synth_points_device(h, 0x0BADC0DE, 0, n, dp) == 0
        assert lib.msm_synth_scalars_device(h, 0x0BADC0DE, 0, L, ds) == 0
        t1 = time.time()
        bh = ctypes.c_void_p()
        assert lib.msm_bases_from_device(h, dp, n, ctypes.byref(bh)) == 0
        lib.msm_device_free(h, dp)
        pre = int(os.environ.get("PRECOMPUTE", "-1"))
        if pre >= 0:
            t2 = time.time()
            assert lib.msm_bases_precompute(h, bh, pre) == 0, lib.msm_last_error(h)
            print("precompute: c=%d in %.3f s" % (lib.msm_bases_table_window(bh), time.time() - t2))
        if os.environ.get("PRECOMPUTE_CHUNKED"):
            t2 = time.time()
            assert lib.msm_bases_precompute_chunked(h, bh, L // chunks) == 0, lib.msm_last_error(h)
            print("precompute_chunked(%d): c=%d in %.3f s" % (L // chunks, lib.msm_bases_table_window(bh), time.time() - t2))
        for c in windows:
            ws.set_window_bits(c)
            best = None
            for it in range(4):
                rc = lib.msm_multiple_multiexp_device(h, bh, ds, L, chunks, do)
                assert rc == 0, (rc, lib.msm_last_error(h))
                t = ws.timings()
                if it > 0 and (best is None or t["total_ms"] < best["total_ms"]):
                    best = t
            pts = n / (best["total_ms"] * 1e-3)
            macs = {0: 21760, 1: 48000, 2: 16 * 10 * 400, 3: 16 * 10 * 888}[curve]  # G2: an Fq2 product = 2 (3 N^2 + N) MACs
            print(json.dumps({"log_L": lg, "lines": lines, "chunks": chunks, "log_n": lg, "c": best["window_bits"], "W": best["num_windows"],
                              "total_ms": round(best["total_ms"], 3), "sort_ms": round(best["sort_ms"], 3),
                              "acc_ms": round(best["accumulate_ms"], 3), "red_ms": round(best["reduce_ms"], 3),
                              "points_per_s": round(pts), "imad_roofline_frac": round(pts * macs / 1.8612e13, 4),
                              "synth_s": round(t1 - t0, 2)}))
        ws.set_window_bits(0)
        lib.msm_bases_free(bh)
        lib.msm_device_free(h, ds)
        lib.msm_device_free(h, do)


if __name__ == "__main__":
    main()
