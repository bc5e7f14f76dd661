#!/usr/bin/env python3
"""Device timing of msm_scalar_fft_device (development aid)."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ec_gpu_b200 as m  # noqa: E402

R = {0: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
     1: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001}
GEN = {0: 5, 1: 7}


def main():
    lib = m.load_library()
    curve = int(os.environ.get("CURVE", "0"))
    ws = m.Workspace(curve)
    h = ws.handle
    for lg in [int(x) for x in (sys.argv[1:] or ["16", "20", "24"])]:
        n = 1 << lg
        r = R[curve]
        om = np.frombuffer((pow(GEN[curve], (r - 1) // n, r) * (1 << 256) % r).to_bytes(32, "little"), dtype=np.uint8).copy()
        d = ctypes.c_void_p()
        assert lib.msm_device_alloc(h, n * 32, ctypes.byref(d)) == 0
        assert lib.msm_synth_scalars_device(h, 5, 0, n, d) == 0
        best = None
        for it in range(4):
            rc = lib.msm_scalar_fft_device(h, d, lg, om.ctypes.data)
            assert rc == 0, lib.msm_last_error(h)
            t = ws.timings()["total_ms"]
            best = t if best is None or t < best else best
        mults = (n // 2) * lg
        print(json.dumps({"curve": curve, "log_n": lg, "device_ms": round(best, 3), "elements_per_s": round(n / (best * 1e-3)),
                          "butterflies_per_s": round(mults / (best * 1e-3)),
                          "hbm_GBps_algorithmic_one_pass": round(n * 64 / (best * 1e-3) / 1e9, 1)}))
        lib.msm_device_free(h, d)


if __name__ == "__main__":
    main()
