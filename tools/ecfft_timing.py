#!/usr/bin/env python3
"""Device timing of msm_ec_fft_device (development aid): the reference bench shape is n = 2^0 .. 2^11
(ag-cuda-ec/benches/ec_fft.rs:24-56)."""
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
    fq = m.fq_bytes(curve)
    for lg in [int(x) for x in (sys.argv[1:] or ["8", "11", "14", "16"])]:
        n = 1 << lg
        r = R[curve]
        omega = pow(GEN[curve], (r - 1) // n, r)
        om = np.zeros((32, 32), dtype=np.uint8)
        for i in range(32):
            om[i] = np.frombuffer((pow(omega, 1 << i, r) * (1 << 256) % r).to_bytes(32, "little"), dtype=np.uint8)
        d_aff, d_jac = ctypes.c_void_p(), ctypes.c_void_p()
        assert lib.msm_device_alloc(h, n * 2 * fq, ctypes.byref(d_aff)) == 0
        assert lib.msm_device_alloc(h, n * 3 * fq, ctypes.byref(d_jac)) == 0
        assert lib.msm_synth_points_device(h, 7, 0, n, d_aff) == 0
        aff = np.zeros((n, 2 * fq), dtype=np.uint8)
        assert lib.msm_memcpy_d2h(h, aff.ctypes.data, d_aff, aff.nbytes) == 0
        jac = np.zeros((n, 3 * fq), dtype=np.uint8)
        jac[:, :2 * fq] = aff
        one = {0: 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d,
               1: 0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd}[curve]
        jac[:, 2 * fq:] = np.frombuffer(one.to_bytes(fq, "little"), dtype=np.uint8)
        best = None
        for it in range(3):
            assert lib.msm_memcpy_h2d(h, d_jac, jac.ctypes.data, jac.nbytes) == 0
            rc = lib.msm_ec_fft_device(h, d_jac, lg, om.ctypes.data, 32)
            assert rc == 0, lib.msm_last_error(h)
            t = ws.timings()["total_ms"]
            best = t if best is None or t < best else best
        butterflies = (n // 2) * lg
        print(json.dumps({"curve": curve, "log_n": lg, "device_ms": round(best, 3), "butterflies": butterflies,
                          "scalar_muls_per_s": round(butterflies / (best * 1e-3)) if best else None}))
        lib.msm_device_free(h, d_aff)
        lib.msm_device_free(h, d_jac)


if __name__ == "__main__":
    main()
