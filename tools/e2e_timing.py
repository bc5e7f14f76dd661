#!/usr/bin/env python3
"""End-to-end (host scalars in, host point out) timing of msm_multiple_multiexp for several
upload pipeline depths (development aid; bench.py is the contract).  Depth 0 = the engine's own choice (which adapts
to the upload rate and device time it measured on the earlier calls of the loop)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ec_gpu_b200 as m  # noqa: E402


def main():
    import torch
    lib = m.load_library()
    curve = int(os.environ.get("CURVE", "0"))
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    depths = [int(x) for x in (sys.argv[2:] or ["1", "2", "4", "8"])]
    L = 1 << lg
    fq = m.fq_bytes(curve)
    ws = m.Workspace(curve)
    h = ws.handle
    dev = torch.device("cuda", 0)
    d_pts = torch.empty(L * 2 * fq, dtype=torch.uint8, device=dev)
    d_sc = torch.empty(L * 32, dtype=torch.uint8, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    assert lib.msm_synth_points_device(h, 0x0BADC0DE, 0, L, p(d_pts)) == 0
    assert lib.msm_synth_scalars_device(h, 0x0BADC0DE, 0, L, p(d_sc)) == 0
    bh = ctypes.c_void_p()
    assert lib.msm_bases_from_device(h, p(d_pts), L, ctypes.byref(bh)) == 0
    del d_pts
    if not os.environ.get("NO_TABLE"):
        assert lib.msm_bases_precompute(h, bh, 0) == 0
    h_sc = torch.empty(L * 32, dtype=torch.uint8, pin_memory=True)
    h_sc.copy_(d_sc)
    h_out = torch.zeros(3 * fq, dtype=torch.uint8, pin_memory=True)
    ref = None
    for depth in depths:
        os.environ["MSM_B200_PIPELINE"] = str(depth)
        best = 1e9
        for it in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = lib.msm_multiple_multiexp(h, bh, ctypes.c_void_p(h_sc.data_ptr()), L, 1, 8, 1, ctypes.c_void_p(h_out.data_ptr()))
            assert rc == 0, lib.msm_last_error(h)
            dt = (time.perf_counter() - t0) * 1e3
            if it > 0:
                best = min(best, dt)
        xy = np.zeros(2 * fq, dtype=np.uint8)
        inf = np.zeros(1, dtype=np.uint8)
        assert lib.msm_to_affine(h, ctypes.c_void_p(h_out.data_ptr()), 1, 0, xy.ctypes.data_as(ctypes.c_void_p),
                                 inf.ctypes.data_as(ctypes.c_void_p)) == 0
        aff = xy.tobytes()
        if ref is None:
            ref = aff
        t = ws.timings()
        print(json.dumps({"log_n": lg, "pipeline": depth, "e2e_ms": round(best, 3), "points_per_s": round(L / best * 1e3),
                          "device_total_ms": round(t["total_ms"], 3), "h2d_ms": round(t["h2d_ms"], 3),
                          "sub_batches": t["sub_batches"],
                          "same_result": aff == ref}))


if __name__ == "__main__":
    main()
