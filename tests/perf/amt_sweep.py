#!/usr/bin/env python3
"""tests/perf/amt_sweep.py -- the AMT bench shape of the reference (ag-cuda-ec/benches/amt.rs:18-55): 10 base lines x 2^21
points sharing one scalar row, cut into 2^7 .. 2^11 groups per line.  MEASUREMENT INFRASTRUCTURE (uses the oracle's
goldens as the checker).

For every group count: this engine through upload -> multiple_multiexp (first call: plain resident copy; second call
builds the window table by policy; then best of 3), and -- for the group counts given on the command line, default 2048 --
the reference's own kernel on the same GPU (oracle/_ref/ref_kernel_bn254, window 8, signed digits as the bench passes
neg_is_cheap = true).  The 2048-group results are compared with tests/golden/fullsize.json.

  python tests/perf/amt_sweep.py [ref_groups ...] > gpurun_out/amt_sweep.jsonl
"""
import ctypes
import hashlib
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
IMAD_PEAK = 148 * 64 * 1.965e9


def digest(jac):
    xy, inf = O.to_affine(0, jac.reshape(-1, 96))
    return hashlib.sha256(xy.tobytes() + inf.tobytes()).hexdigest()


def main():
    ref_groups = [int(x) for x in sys.argv[1:]] or [2048]
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize.json")))["bn254_amt_10x2p21_2048"]
    lines, L = golden["lines"], golden["L"]
    n = lines * L
    lib = m.load_library()
    ws = m.Workspace(0)
    h = ws.handle
    dp, ds = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.msm_device_alloc(h, n * 64, ctypes.byref(dp)) == 0
    assert lib.msm_device_alloc(h, L * 32, ctypes.byref(ds)) == 0
    assert lib.msm_synth_points_device(h, SEED, 0, n, dp) == 0
    assert lib.msm_synth_scalars_device(h, SEED, 0, L, ds) == 0
    pts = np.zeros((n, 64), dtype=np.uint8)
    sc = np.zeros((L, 32), dtype=np.uint8)
    assert lib.msm_memcpy_d2h(h, pts.ctypes.data, dp, pts.nbytes) == 0
    assert lib.msm_memcpy_d2h(h, sc.ctypes.data, ds, sc.nbytes) == 0
    lib.msm_device_free(h, dp)
    tmp = tempfile.mkdtemp(prefix="amt_")
    fb, fe, fo = (os.path.join(tmp, x) for x in ("bases.bin", "exps.bin", "out.bin"))
    if any(g in ref_groups for g in (128, 256, 512, 1024, 2048)) and os.path.exists(EXE):
        pts.tofile(fb)
        sc.tofile(fe)
    for groups in (128, 256, 512, 1024, 2048):
        bases = m.upload_multiexp_bases(ws, pts)
        do = ctypes.c_void_p()
        assert lib.msm_device_alloc(h, lines * groups * 96, ctypes.byref(do)) == 0
        times, tables = [], []
        for _ in range(5):
            assert lib.msm_multiple_multiexp_device(h, bases._h, ds, L, groups, do) == 0, lib.msm_last_error(h)
            times.append(ws.timings()["total_ms"])
            tables.append(int(lib.msm_bases_table_window(bases._h)))
        t = ws.timings()
        out = np.zeros((lines * groups, 96), dtype=np.uint8)
        assert lib.msm_memcpy_d2h(h, out.ctypes.data, do, out.nbytes) == 0
        best = min(times[2:])
        canon = 32 * 10 * 136  # c = 8 canonical count for small per-group MSMs
        rec = {"shape": "AMT 10 lines x 2^21, %d groups of %d points" % (groups, L // groups), "groups": groups,
               "engine_ms_first_call_plain": round(times[0], 3), "engine_ms": round(best, 3),
               "table_window": tables[-1], "window_bits": t["window_bits"], "num_windows": t["num_windows"],
               "points_per_s": n / (best * 1e-3), "imad_roofline_frac_c8_count": n / (best * 1e-3) * canon / IMAD_PEAK,
               "phases_ms": {k: round(t[k], 3) for k in ("sort_ms", "accumulate_ms", "reduce_ms")}}
        if groups == golden["num_chunks"]:
            rec["result_matches_golden"] = digest(out) == golden["sha256"]
        if groups in ref_groups and os.path.exists(EXE):
            line = subprocess.run([EXE, fb, fe, str(L), str(groups), "8", "1", "2", fo], check=True, capture_output=True,
                                  text=True).stdout.strip().splitlines()[-1]
            ref = json.loads(line)
            rout = np.fromfile(fo, dtype=np.uint8).reshape(-1, 96)
            rec["reference_kernel_ms"] = ref["ms_best"]
            rec["reference_kernel_equals_engine"] = digest(rout) == digest(out)
            rec["speedup_vs_reference_kernel"] = round(ref["ms_best"] / best, 1)
        lib.msm_device_free(h, do)
        bases.free()
        print(json.dumps(rec), flush=True)
    lib.msm_device_free(h, ds)


if __name__ == "__main__":
    main()
