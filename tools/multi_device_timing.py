#!/usr/bin/env python3
"""In-library multi-device path (msm_ctx over several GPUs, one process): MultiexpKernel::multiexp semantics
(ec-gpu-proxy/src/multiexp.rs:324-400) -- contiguous shards, one host thread per device, partial points gathered on
device 0 over peer copies and summed there.  Times msm_multiexp_resident (pinned host scalars in, host point out) at
1, 2, 4, 8 devices for BN254 2^24 and checks every result against tests/golden/fullsize.json.

  python tools/multi_device_timing.py [log_n] > gpurun_out/multi_device.jsonl      (needs the GPUs of one box)
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ec_gpu_b200 as m  # noqa: E402

SEED = 0x0BADC0DE


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << log_n
    lib = m.load_library()
    ndev = lib.msm_device_count()
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize.json"))).get("bn254_2p%d" % log_n)
    # inputs: generated on device 0, copied to (pinned) host memory
    ws0 = m.Workspace(0)
    h0 = ws0.handle
    dp, ds = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.msm_device_alloc(h0, n * 64, ctypes.byref(dp)) == 0
    assert lib.msm_device_alloc(h0, n * 32, ctypes.byref(ds)) == 0
    assert lib.msm_synth_points_device(h0, SEED, 0, n, dp) == 0
    assert lib.msm_synth_scalars_device(h0, SEED, 0, n, ds) == 0
    pts = np.zeros((n, 64), dtype=np.uint8)
    sc = np.zeros((n, 32), dtype=np.uint8)
    assert lib.msm_memcpy_d2h(h0, pts.ctypes.data, dp, pts.nbytes) == 0
    assert lib.msm_memcpy_d2h(h0, sc.ctypes.data, ds, sc.nbytes) == 0
    lib.msm_device_free(h0, dp)
    lib.msm_device_free(h0, ds)
    assert lib.msm_host_register(sc.ctypes.data, sc.nbytes) == 0
    for k in (1, 2, 4, 8):
        if k > ndev:
            break
        kern = m.MultiexpKernel.create(list(range(k)), 0)
        res = kern.upload_bases(pts)
        times = []
        for it in range(7):
            t0 = time.perf_counter()
            out = kern.multiexp_resident(res, sc, 0)
            times.append((time.perf_counter() - t0) * 1e3)
        aff = np.zeros(64, dtype=np.uint8)
        inf = np.zeros(1, dtype=np.uint8)
        assert lib.msm_to_affine(kern.workspace.handle, out.ctypes.data, 1, 0, aff.ctypes.data, inf.ctypes.data) == 0
        ok = None
        if golden:
            ok = (bytes(aff[:32][::-1]).hex() == golden["result"]["x"] and bytes(aff[32:][::-1]).hex() == golden["result"]["y"])
        t = kern.workspace.timings()
        print(json.dumps({"devices": k, "log_n": log_n, "call": "msm_multiexp_resident (host scalars, host result)",
                          "ms_first_call_plain": round(times[0], 3), "ms_second_call_builds_tables": round(times[1], 3),
                          "ms_best": round(min(times[2:]), 3), "ms_median": round(sorted(times[2:])[len(times[2:]) // 2], 3),
                          "points_per_s": n / (min(times[2:]) * 1e-3), "dev0_window_bits": t["window_bits"],
                          "dev0_phases_ms": {x: round(t[x], 3) for x in ("h2d_ms", "sort_ms", "accumulate_ms", "reduce_ms", "total_ms")},
                          "result_matches_golden": ok}), flush=True)
        res.free()
        kern.workspace.close()
    lib.msm_host_unregister(sc.ctypes.data)


if __name__ == "__main__":
    main()
