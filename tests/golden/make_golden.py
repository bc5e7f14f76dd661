#!/usr/bin/env python3
"""Generates the committed golden fixtures of tests/golden/.

  pyref_vectors.json   from oracle/pyref.py (independent pure-Python big-integer arithmetic):
                       field products, curve KATs, small MSMs incl. edge cases -- needs nothing
                       but Python.
  ref_cl_vectors.json  from oracle/_ref (the REFERENCE'S OWN device sources ag-build/cl/*.cl,
                       compiled for the host by oracle/build_ref.py and run in this container):
                       FIELD_add/sub/mul/sqr/double/mont/unmont, POINT_add/add_mixed/double and the
                       POINT_multiexp kernel.  Needs /root/reference; the fixture travels instead.

  ec_fft_vectors.json  from oracle/pyref.py: the DFT of G1 points evaluated by its definition.
  fr_fft_vectors.json  the DFT over the scalar field evaluated by its definition (Python integers).
  ref_cl_fft_vectors.json  outputs of the reference's own fft.cl / ec-fft.cl kernels run on the host (needs /root/reference).
  ref_cl_g2_vectors.json   the reference's field2.cl / ec.cl / multiexp.cl instantiated over Fq2, run on the host.
  g2_vectors.json      G2 over Fq2 from oracle/pyref.py's G2Params (--only-g2 regenerates just this file).

Run from the repo root:  python tests/golden/make_golden.py   (--only-fft: just the FFT files)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0x0BADC0DE


def hx(b):
    return bytes(b).hex()


def make_pyref():
    out = {"generator": "tests/golden/make_golden.py (oracle/pyref.py)", "curves": {}}
    for cv in (P.BN254, P.BLS12_381):
        rng = np.random.default_rng(2024 + cv.curve_id)
        c = {}
        # field products in the API Montgomery domain: mont(a)*mont(b)/R = mont(a*b)
        fa = [int.from_bytes(rng.bytes(cv.fq_bytes + 8), "little") % cv.p for _ in range(16)]
        fb = [int.from_bytes(rng.bytes(cv.fq_bytes + 8), "little") % cv.p for _ in range(16)]
        fa[:3], fb[:3] = [0, 1, cv.p - 1], [cv.p - 1, cv.p - 1, cv.p - 1]
        c["fq_mul"] = {
            "a_mont": [hx(cv.fq_to_bytes(cv.to_mont(x))) for x in fa],
            "b_mont": [hx(cv.fq_to_bytes(cv.to_mont(x))) for x in fb],
            "ab_mont": [hx(cv.fq_to_bytes(cv.to_mont(x * y % cv.p))) for x, y in zip(fa, fb)],
            "a_plus_b_mont": [hx(cv.fq_to_bytes(cv.to_mont((x + y) % cv.p))) for x, y in zip(fa, fb)],
            "a_minus_b_mont": [hx(cv.fq_to_bytes(cv.to_mont((x - y) % cv.p))) for x, y in zip(fa, fb)],
        }
        # multiples of the generator (canonical affine)
        c["kG"] = {str(k): [hex(v) for v in cv.mul(k, cv.g)] for k in (1, 2, 3, 7, cv.r - 1)}
        # synthetic-input generator prefix
        n = 40
        sc = P.gen_scalars(cv, SEED, 0, n)
        pts = P.gen_points(cv, SEED, 0, n)
        c["synthetic"] = {"seed": SEED, "scalars": hx(P.scalars_to_bytes(sc)), "points_mont": hx(P.points_to_bytes(cv, pts))}
        # MSMs: plain, and with edge cases
        cases = []
        cases.append(("random40", sc, pts))
        esc, epts = list(sc), list(pts)
        esc[0], esc[1], esc[2] = 0, 1, cv.r - 1
        epts[3] = None                      # identity base
        epts[5], esc[5] = epts[4], esc[4]   # repeated (point, scalar)
        epts[7], esc[7] = cv.neg(epts[6]), esc[6]  # P and -P with equal scalars
        esc[8] = (1 << 128) - 1
        esc[9] = 0x8000
        esc[10] = 0xFFFF
        cases.append(("edge40", esc, epts))
        cases.append(("single", [sc[0]], [pts[0]]))
        cases.append(("all_zero_scalars", [0] * 8, pts[:8]))
        c["msm"] = []
        for name, s, p in cases:
            res = cv.msm(s, p)
            c["msm"].append({
                "name": name,
                "scalars": hx(P.scalars_to_bytes(s)),
                "points_mont": hx(P.points_to_bytes(cv, p)),
                "result_affine": None if res is None else [hex(res[0]), hex(res[1])],
            })
        # batched shape: 2 lines x 4 chunks x 8 points (ag_cuda_ec::multiple_multiexp semantics)
        L, lines, chunks = 32, 2, 4
        bsc = P.gen_scalars(cv, SEED + 1, 0, L)
        bpts = P.gen_points(cv, SEED + 1, 0, L * lines)
        res = []
        cl = L // chunks
        for line in range(lines):
            for ch in range(chunks):
                r = cv.msm(bsc[ch * cl:(ch + 1) * cl], bpts[line * L + ch * cl: line * L + (ch + 1) * cl])
                res.append(None if r is None else [hex(r[0]), hex(r[1])])
        c["multiple_multiexp"] = {"L": L, "lines": lines, "chunks": chunks, "scalars": hx(P.scalars_to_bytes(bsc)),
                                  "points_mont": hx(P.points_to_bytes(cv, bpts)), "results_affine": res}
        out["curves"][cv.name] = c
    with open(os.path.join(HERE, "pyref_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


def make_ref_cl():
    from oracle import oracle as O
    from oracle import ref_cl as R

    if not os.path.isdir("/root/reference/ag-build/cl"):
        print("reference sources absent: ref_cl_vectors.json not regenerated")
        return
    out = {"generator": "tests/golden/make_golden.py (oracle/_ref = /root/reference/ag-build/cl/*.cl built for the host)",
           "curves": {}}
    for curve, name in ((0, "bn254"), (1, "bls12_381")):
        fq = O.FQ_BYTES[curve]
        rng = np.random.default_rng(77 + curve)
        p = int.from_bytes(O.constant(curve, 0).tobytes(), "little")
        n = 12
        vals_a = [int.from_bytes(rng.bytes(fq + 8), "little") % p for _ in range(n)]
        vals_b = [int.from_bytes(rng.bytes(fq + 8), "little") % p for _ in range(n)]
        vals_a[:2], vals_b[:2] = [0, p - 1], [p - 1, p - 1]
        a = np.frombuffer(b"".join(v.to_bytes(fq, "little") for v in vals_a), dtype=np.uint8).copy()
        b = np.frombuffer(b"".join(v.to_bytes(fq, "little") for v in vals_b), dtype=np.uint8).copy()
        c = {"fq": {"a": hx(a), "b": hx(b), "ops": {}}}
        for op, nm in enumerate(["add", "sub", "mul", "sqr", "double", "mont", "unmont"]):
            c["fq"]["ops"][nm] = hx(R.fq_op(curve, op, a, b))
        # curve ops on Jacobian inputs produced by the reference's own arithmetic (k*P by repeated ops)
        pts = O.gen_points(curve, 11, n)
        one = O.constant(curve, 1)
        lifted = np.zeros((n, 3 * fq), dtype=np.uint8)
        lifted[:, :2 * fq] = pts
        lifted[:, 2 * fq:] = one
        dbl = R.ec_op(curve, 2, lifted)
        trip = R.ec_op(curve, 1, dbl, pts)
        c["ec"] = {"affine": hx(pts), "double": hx(dbl), "double_plus_affine": hx(trip),
                   "add_double_triple": hx(R.ec_op(curve, 0, dbl, trip))}
        # POINT_multiexp kernel: 2 lines x 4 chunks x 16 points
        L, lines, chunks = 64, 2, 4
        sc = O.gen_scalars(curve, SEED + 2, L)
        bp = O.gen_points(curve, SEED + 2, L * lines)
        c["multiexp"] = {"L": L, "lines": lines, "chunks": chunks, "scalars": hx(sc), "points_mont": hx(bp), "runs": []}
        for w, neg in ((3, True), (4, False), (7, True)):
            c["multiexp"]["runs"].append({"window_size": w, "neg_is_cheap": neg,
                                          "results_jacobian": hx(R.multiple_multiexp(curve, bp, sc, chunks, w, neg))})
        out["curves"][name] = c
    with open(os.path.join(HERE, "ref_cl_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


def make_ec_fft():
    """ec_fft_vectors.json: the DFT of G1 points by its definition, out[k] = sum_j omega^(j k) P_j, in
    oracle/pyref.py's big-integer affine arithmetic (no FFT structure at all).  omega = g^((r-1)/n) with
    g the multiplicative generator arkworks uses for Fr (5 for BN254, 7 for BLS12-381), so omega is
    Scalar::get_root_of_unity(n) of ag-cuda-ec/src/ec_fft.rs:120."""
    out = {"generator": "tests/golden/make_golden.py make_ec_fft (oracle/pyref.py, naive O(n^2) DFT)", "curves": {}}
    for cv, g in ((P.BN254, 5), (P.BLS12_381, 7)):
        cases = []
        for log_n in (3, 4):
            n = 1 << log_n
            omega = pow(g, (cv.r - 1) // n, cv.r)
            assert pow(omega, n, cv.r) == 1 and pow(omega, n // 2, cv.r) == cv.r - 1
            pts = P.gen_points(cv, SEED + 3 + log_n, 0, n)
            pts[3] = None                      # the identity as an input
            pts[5] = pts[4]                    # equal neighbours: a butterfly hits the doubling branch
            lam = {2: 7, 6: cv.p - 5}          # non-trivial Jacobian representatives (x l^2, y l^3, l)
            jac = b""
            for i, pt in enumerate(pts):
                if pt is None:
                    x, y, z = 0, 1, 0
                else:
                    l = lam.get(i, 1)
                    x, y, z = pt[0] * l * l % cv.p, pt[1] * l * l * l % cv.p, l
                jac += cv.fq_to_bytes(cv.to_mont(x)) + cv.fq_to_bytes(cv.to_mont(y)) + cv.fq_to_bytes(cv.to_mont(z))
            R_fr = 1 << 256
            omegas = b"".join((pow(omega, 1 << i, cv.r) * R_fr % cv.r).to_bytes(32, "little") for i in range(32))
            res = []
            for k in range(n):
                acc = None
                for j, pt in enumerate(pts):
                    if pt is not None:
                        acc = cv.add(acc, cv.mul(pow(omega, j * k % n, cv.r), pt))
                res.append(None if acc is None else [hex(acc[0]), hex(acc[1])])
            cases.append({"log_n": log_n, "omega": hex(omega), "omegas_mont": hx(omegas), "input_jacobian": hx(jac),
                          "output_affine": res})
        out["curves"][cv.name] = cases
    with open(os.path.join(HERE, "ec_fft_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


def make_fr_fft():
    """fr_fft_vectors.json: the DFT over the scalar field by its definition in Python integers."""
    out = {"generator": "tests/golden/make_golden.py make_fr_fft (Python integers, naive O(n^2) DFT)", "curves": {}}
    R = 1 << 256
    for cv, g in ((P.BN254, 5), (P.BLS12_381, 7)):
        rng = np.random.default_rng(4242 + cv.curve_id)
        cases = []
        for log_n in (1, 4, 7):
            n = 1 << log_n
            omega = pow(g, (cv.r - 1) // n, cv.r)
            a = [int.from_bytes(rng.bytes(40), "little") % cv.r for _ in range(n)]
            a[0], a[-1] = 0, cv.r - 1
            res = [sum(a[j] * pow(omega, j * k % n, cv.r) for j in range(n)) % cv.r for k in range(n)]
            cases.append({"log_n": log_n, "omega_mont": hx((omega * R % cv.r).to_bytes(32, "little")),
                          "input_mont": hx(b"".join((x * R % cv.r).to_bytes(32, "little") for x in a)),
                          "output_mont": hx(b"".join((x * R % cv.r).to_bytes(32, "little") for x in res))})
        out["curves"][cv.name] = cases
    with open(os.path.join(HERE, "fr_fft_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


def make_g2():
    """g2_vectors.json: G2 over Fq2 from oracle/pyref.py's G2Params (independent big-integer affine
    arithmetic): Fq2 products, multiples of the generator, the synthetic-input prefix, small MSMs with
    the edge cases, and one batched shape."""
    out = {"generator": "tests/golden/make_golden.py make_g2 (oracle/pyref.py G2Params)", "curves": {}}
    for cv in (P.BN254_G2, P.BLS12_381_G2):
        rng = np.random.default_rng(3030 + cv.curve_id)
        rnd = lambda: (int.from_bytes(rng.bytes(cv.base_bytes + 8), "little") % cv.p,  # noqa: E731
                       int.from_bytes(rng.bytes(cv.base_bytes + 8), "little") % cv.p)
        c = {}
        fa, fb = [rnd() for _ in range(8)], [rnd() for _ in range(8)]
        fa[:2], fb[:2] = [(0, 0), (cv.p - 1, 1)], [(5, 7), (cv.p - 1, cv.p - 1)]
        c["fq2"] = {"a_mont": [hx(cv.fq_to_bytes(cv.to_mont(x))) for x in fa],
                    "b_mont": [hx(cv.fq_to_bytes(cv.to_mont(x))) for x in fb],
                    "ab_mont": [hx(cv.fq_to_bytes(cv.to_mont(cv.f2mul(x, y)))) for x, y in zip(fa, fb)],
                    "a_inv_mont": [hx(cv.fq_to_bytes(cv.to_mont(cv.f2inv(x)))) for x in fa[1:]]}
        flat = lambda pt: None if pt is None else [[hex(v) for v in pt[0]], [hex(v) for v in pt[1]]]  # noqa: E731
        c["kG"] = {str(k): flat(cv.mul(k, cv.g)) for k in (1, 2, 3, 7, cv.r - 1)}
        n = 12
        sc = P.gen_scalars(cv, SEED, 0, n)
        pts = P.gen_points(cv, SEED, 0, n)
        c["synthetic"] = {"seed": SEED, "scalars": hx(P.scalars_to_bytes(sc)), "points_mont": hx(P.points_to_bytes(cv, pts))}
        esc, epts = list(sc), list(pts)
        esc[0], esc[1], esc[2] = 0, 1, cv.r - 1
        epts[3] = None
        epts[5], esc[5] = epts[4], esc[4]
        epts[7], esc[7] = cv.neg(epts[6]), esc[6]
        esc[8] = (1 << 128) - 1
        c["msm"] = []
        for name, s_, p_ in (("random12", sc, pts), ("edge12", esc, epts)):
            c["msm"].append({"name": name, "scalars": hx(P.scalars_to_bytes(s_)), "points_mont": hx(P.points_to_bytes(cv, p_)),
                             "result_affine": flat(cv.msm(s_, p_))})
        L, lines, chunks = 8, 2, 2
        res = []
        for line in range(lines):
            for ch in range(chunks):
                lo = ch * (L // chunks)
                res.append(flat(cv.msm(sc[lo:lo + L // chunks], pts[(line * L + lo) % n:][: L // chunks])))
        bpts = [pts[(i % n)] for i in range(L * lines)]
        res = []
        for line in range(lines):
            for ch in range(chunks):
                lo = ch * (L // chunks)
                res.append(flat(cv.msm(sc[lo:lo + L // chunks], bpts[line * L + lo: line * L + lo + L // chunks])))
        c["multiple_multiexp"] = {"L": L, "lines": lines, "chunks": chunks, "scalars": hx(P.scalars_to_bytes(sc[:L])),
                                  "points_mont": hx(P.points_to_bytes(cv, bpts)), "results_affine": res}
        out["curves"][cv.name] = c
    with open(os.path.join(HERE, "g2_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


def make_ref_cl_fft():
    """ref_cl_fft_vectors.json: outputs of the REFERENCE'S OWN FFT kernels (ag-build/cl/fft.cl,
    ag-build/cl/ec-fft.cl) executed on the host by oracle/build_ref.py under restatements of their host
    pass loops (ec-gpu-proxy/src/fft.rs:50-136, ag-cuda-ec/src/ec_fft.rs:13-99).  Needs /root/reference."""
    from oracle import oracle as O
    from oracle import ref_cl as R

    if not os.path.isdir("/root/reference/ag-build/cl"):
        print("reference sources absent: ref_cl_fft_vectors.json not regenerated")
        return
    out = {"generator": "tests/golden/make_golden.py make_ref_cl_fft (oracle/_ref: the reference's fft.cl / ec-fft.cl on the host)",
           "curves": {}}
    for curve, name, g in ((0, "bn254", 5), (1, "bls12_381", 7)):
        cv = P.CURVES[curve]
        fq = O.FQ_BYTES[curve]
        c = {"fr_fft": [], "ec_fft": []}
        for log_n in (3, 9):   # 9 > MAX_LOG2_RADIX = 8: two passes of the reference kernel
            n = 1 << log_n
            omega = pow(g, (cv.r - 1) // n, cv.r)
            om = np.frombuffer((omega * (1 << 256) % cv.r).to_bytes(32, "little"), dtype=np.uint8).copy()
            a = O.gen_scalars(curve, SEED + log_n, n)
            c["fr_fft"].append({"log_n": log_n, "omega_mont": hx(om), "input_mont": hx(a), "output_mont": hx(R.fr_fft(curve, a, om))})
        for log_n in (2, 5):
            n = 1 << log_n
            omega = pow(g, (cv.r - 1) // n, cv.r)
            oms = np.zeros((32, 32), dtype=np.uint8)
            for i in range(32):
                oms[i] = np.frombuffer((pow(omega, 1 << i, cv.r) * (1 << 256) % cv.r).to_bytes(32, "little"), dtype=np.uint8)
            jac = np.zeros((n, 3 * fq), dtype=np.uint8)
            jac[:, :2 * fq] = O.gen_points(curve, SEED + 7 + log_n, n)
            jac[:, 2 * fq:] = O.constant(curve, 1)
            res = R.ec_fft(curve, jac, oms)
            xy, inf = O.to_affine(curve, res)  # the Jacobian representative is the reference's; the fixture keeps the group elements
            c["ec_fft"].append({"log_n": log_n, "omegas_mont": hx(oms), "input_jacobian": hx(jac),
                                "output_affine_canonical": hx(xy), "output_is_inf": [int(v) for v in inf]})
        out["curves"][name] = c
    with open(os.path.join(HERE, "ref_cl_fft_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


def make_ref_cl_g2():
    """ref_cl_g2_vectors.json: outputs of the REFERENCE'S OWN field2.cl / ec.cl / multiexp.cl instantiated over
    Fq2 exactly as its SourceBuilder would (ag-build/src/source/synthesis.rs:100-110) and run on the host by
    oracle/build_ref.py: FIELD2 products, G2 double / add / mixed add, and the POINT_multiexp kernel."""
    from oracle import oracle as O
    from oracle import ref_cl as R

    if not os.path.isdir("/root/reference/ag-build/cl"):
        print("reference sources absent: ref_cl_g2_vectors.json not regenerated")
        return
    out = {"generator": "tests/golden/make_golden.py make_ref_cl_g2 (oracle/_ref: the reference's field2.cl / ec.cl / multiexp.cl over Fq2)",
           "curves": {}}
    for g1, g2, name in ((0, 2, "bn254_g2"), (1, 3, "bls12_381_g2")):
        fq = O.FQ_BYTES[g2]
        rng = np.random.default_rng(500 + g2)
        p = int.from_bytes(O.constant(g2, 0)[: fq // 2].tobytes(), "little")
        n = 8
        mk = lambda: np.frombuffer(b"".join((int.from_bytes(rng.bytes(fq), "little") % p).to_bytes(fq // 2, "little")  # noqa: E731
                                            for _ in range(2 * n)), dtype=np.uint8).reshape(n, fq).copy()
        a, b = mk(), mk()
        c = {"fq2": {"a": hx(a), "b": hx(b), "ops": {nm: hx(R.fq2_op(g1, op, a, b))
                                                      for op, nm in enumerate(["add", "sub", "mul", "sqr", "double"])}}}
        pts = O.gen_points(g2, 21, 2 * n)
        lifted = np.zeros((n, 3 * fq), dtype=np.uint8)
        lifted[:, :2 * fq] = pts[:n]
        lifted[:, 2 * fq:] = O.constant(g2, 1)
        dbl = R.g2_ec_op(g1, 2, lifted)
        c["ec"] = {"affine": hx(pts), "double": hx(dbl), "double_plus_affine": hx(R.g2_ec_op(g1, 1, dbl, pts[n:]))}
        L, lines, chunks = 32, 2, 4
        sc = O.gen_scalars(g2, SEED + 5, L)
        bp = O.gen_points(g2, SEED + 5, L * lines)
        c["multiexp"] = {"L": L, "lines": lines, "chunks": chunks, "scalars": hx(sc), "points_mont": hx(bp), "runs": []}
        for w, neg in ((3, True), (5, False)):
            res = R.g2_multiple_multiexp(g1, bp, sc, chunks, w, neg)
            xy, inf = O.to_affine(g2, res)
            c["multiexp"]["runs"].append({"window_size": w, "neg_is_cheap": neg, "results_affine_canonical": hx(xy),
                                          "results_is_inf": [int(v) for v in inf]})
        out["curves"][name] = c
    with open(os.path.join(HERE, "ref_cl_g2_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if "--only-ec-fft" not in sys.argv and "--only-fft" not in sys.argv and "--only-g2" not in sys.argv:
        make_pyref()
        make_ref_cl()
    if "--only-g2" not in sys.argv:
        make_ec_fft()
        make_fr_fft()
        make_ref_cl_fft()
    if "--only-fft" not in sys.argv:
        make_g2()
        make_ref_cl_g2()
    for fn in ("pyref_vectors.json", "ref_cl_vectors.json", "ec_fft_vectors.json", "fr_fft_vectors.json", "ref_cl_fft_vectors.json", "g2_vectors.json", "ref_cl_g2_vectors.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")
