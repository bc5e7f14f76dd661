#!/usr/bin/env python3
"""tests/perf/ref_kernel_b200.py -- the kernel this engine replaces, timed on the same B200 (SURVEY.md section 8d,
"second baseline").  TEST / MEASUREMENT INFRASTRUCTURE.

oracle/_ref/ref_kernel_bn254 is the reference's own `POINT_multiexp` (ag-build/cl/multiexp.cl:217-264, instantiated for
BN254 G1 by oracle/build_ref.py exactly as SourceBuilder does, compiled for sm_100a) under a small driver that launches it
with the geometry of ag_cuda_ec::multiple_multiexp (ag-cuda-ec/src/multiexp.rs:27-72).  Configurations:

  A  the reference's own bench, ag-cuda-ec/benches/multiexp.rs:19-22,56: 2^22 points, 1024 chunks, window 8, unsigned
  B  2^24 points (BASELINE.json configs[2]) with the same 1024 chunks / window 8; the caller sums the 1024 partials
  C  as A with signed digits (neg_is_cheap = true), the AMT bench's setting (ag-cuda-ec/benches/amt.rs:37-55)

Results are checked against tests/golden/fullsize.json (A, C: the 1024 results; B: their sum), and the engine's time for
the same call is printed next to it.  Usage (GPU box):  python tests/perf/ref_kernel_b200.py > gpurun_out/ref_kernel.jsonl
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ec_gpu_b200 as m  # noqa: E402
from oracle import oracle as O  # noqa: E402

SEED = 0x0BADC0DE
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_kernel_bn254")


def hex_points(jac):
    xy, inf = O.to_affine(0, jac.reshape(-1, 96))
    return [{"x": bytes(r[:32][::-1]).hex(), "y": bytes(r[32:][::-1]).hex(), "inf": int(i)} for r, i in zip(xy, inf)]


def main():
    if not os.path.exists(EXE):
        print(json.dumps({"unavailable": "oracle/_ref/ref_kernel_bn254 not built (python oracle/build_ref.py --cuda)"}))
        return
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize.json")))
    lib = m.load_library()
    ws = m.Workspace(0)
    h = ws.handle
    n = 1 << 24
    dp, ds = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.msm_device_alloc(h, n * 64, ctypes.byref(dp)) == 0
    assert lib.msm_device_alloc(h, n * 32, ctypes.byref(ds)) == 0
    assert lib.msm_synth_points_device(h, SEED, 0, n, dp) == 0
    assert lib.msm_synth_scalars_device(h, SEED, 0, n, ds) == 0
    pts = np.zeros((n, 64), dtype=np.uint8)
    sc = np.zeros((n, 32), dtype=np.uint8)
    assert lib.msm_memcpy_d2h(h, pts.ctypes.data, dp, pts.nbytes) == 0
    assert lib.msm_memcpy_d2h(h, sc.ctypes.data, ds, sc.nbytes) == 0
    lib.msm_device_free(h, dp)
    tmp = tempfile.mkdtemp(prefix="refk_")
    for name, log_n, chunks, window, neg in (("A", 22, 1024, 8, 0), ("C", 22, 1024, 8, 1), ("B", 24, 1024, 8, 0)):
        L = 1 << log_n
        fb, fe, fo = (os.path.join(tmp, x) for x in ("bases.bin", "exps.bin", "out.bin"))
        pts[:L].tofile(fb)
        sc[:L].tofile(fe)
        line = subprocess.run([EXE, fb, fe, str(L), str(chunks), str(window), str(neg), "3", fo], check=True,
                              capture_output=True, text=True).stdout.strip().splitlines()[-1]
        rec = json.loads(line)
        out = np.fromfile(fo, dtype=np.uint8).reshape(-1, 96)
        if log_n == 22:
            rec["matches_golden"] = hex_points(out) == golden["bn254_batched_1024x4096"]["results"]
        else:
            acc = out[0:1].copy()
            for i in range(1, out.shape[0]):
                acc = O.ec_op(0, 0, acc, out[i:i + 1].copy())
            rec["matches_golden"] = hex_points(acc)[0] == golden["bn254_2p24"]["result"]
        rec["config"] = name
        rec["points_per_s"] = L / (rec["ms_best"] * 1e-3)
        # this engine, same call (device-resident scalars; plain resident copy, then with the policy's table)
        bases = m.upload_multiexp_bases(ws, pts[:L])
        do = ctypes.c_void_p()
        assert lib.msm_device_alloc(h, chunks * 96, ctypes.byref(do)) == 0
        times = []
        for _ in range(5):
            assert lib.msm_multiple_multiexp_device(h, bases._h, ds, L, chunks, do) == 0
            times.append(ws.timings()["total_ms"])
        rec["engine_ms_first_call_plain"] = round(times[0], 3)
        rec["engine_ms_table"] = round(min(times[2:]), 3)
        rec["engine_table_window"] = int(lib.msm_bases_table_window(bases._h))
        rec["speedup_vs_reference_kernel"] = round(rec["ms_best"] / min(times[2:]), 1)
        lib.msm_device_free(h, do)
        bases.free()
        print(json.dumps(rec), flush=True)
    lib.msm_device_free(h, ds)


if __name__ == "__main__":
    main()
