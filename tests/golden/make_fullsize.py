#!/usr/bin/env python3
"""tests/golden/make_fullsize.py -- full-size known answers for the configurations BASELINE.json quotes.

TEST INFRASTRUCTURE.  Runs the CPU oracle (oracle/msm_oracle.cpp, the restatement of
ec-gpu-proxy/src/multiexp_cpu.rs:244-367 that tests/test_oracle.py pins against the reference's own
kernels) ONCE on the seed-0x0badc0de synthetic workloads of SURVEY.md section 8d and stores the results
as canonical (non-Montgomery) affine coordinates, the parity notion of
ec-gpu-proxy/tests/multiexp.rs:99 (`into_affine()` on both sides):

  bn254_2p20 / 2p21 / 2p22 / 2p23 / 2p24   one MSM over the first 2^k points   (configs[1], configs[2] and its shards)
  bls12_381_2p22                           configs[3]
  bn254_batched_1024x4096                  configs[4]: 1024 MSMs of 2^12 points (ag-cuda-ec/benches/multiexp.rs:19-22,56)
  bn254_amt_10x2p21_2048                   10 lines x 2^21 points, 2048 chunks (ag-cuda-ec/benches/amt.rs:18-55):
                                           SHA-256 of all 20480 results + the first 8 of them
  bls12_381_batched_256x4096               256 BLS12-381 MSMs of 2^12 points (SHA-256 + the first 8 results)
  bn254_g2_2p20, bls12_381_g2_2p18         G2 over Fq2: one MSM each

Usage:  python tests/golden/make_fullsize.py [--only NAME ...]      (about ten minutes on 8 cores)
The output, tests/golden/fullsize.json, is committed; the GPU tests and bench.py compare against it.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SEED = 0x0BADC0DE
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize.json")


def affine_hex(curve, jac):
    xy, inf = O.to_affine(curve, jac)
    fq = O.FQ_BYTES[curve]
    return [{"x": bytes(r[:fq][::-1]).hex(), "y": bytes(r[fq:][::-1]).hex(), "inf": int(i)} for r, i in zip(xy, inf)]


def affine_digest(curve, jac):
    xy, inf = O.to_affine(curve, jac)
    return hashlib.sha256(xy.tobytes() + inf.tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    O.build()
    try:
        with open(OUT) as f:
            doc = json.load(f)
    except (OSError, ValueError):
        doc = {}
    doc["_about"] = ("canonical affine results of the CPU oracle on the seed 0x0badc0de synthetic workloads "
                     "(SURVEY.md section 8d); big-endian hex; made by tests/golden/make_fullsize.py")
    doc["seed"] = SEED
    want = lambda name: args.only is None or name in args.only  # noqa: E731

    # --- BN254: one point stream, nested prefixes
    sizes = [k for k in (20, 21, 22, 23, 24) if want("bn254_2p%d" % k)]
    need_bn = sizes or want("bn254_batched_1024x4096")
    if need_bn:
        n_max = 1 << max(sizes + [22])
        t0 = time.time()
        pts = O.gen_points(0, SEED, n_max)
        sc = O.gen_scalars(0, SEED, n_max)
        print("bn254 inputs 2^%d: %.1f s" % (n_max.bit_length() - 1, time.time() - t0), flush=True)
        for k in sizes:
            t0 = time.time()
            r = O.multiexp_cpu(0, pts[: 1 << k], sc[: 1 << k])
            doc["bn254_2p%d" % k] = {"curve": 0, "n": 1 << k, "result": affine_hex(0, r)[0],
                                     "oracle_seconds": round(time.time() - t0, 2), "host_threads": O.ncores()}
            print("bn254 2^%d: %.1f s" % (k, time.time() - t0), flush=True)
        if want("bn254_batched_1024x4096"):
            t0 = time.time()
            L = 1 << 22
            r = O.multiple_multiexp(0, pts[:L], sc[:L], 1024)
            doc["bn254_batched_1024x4096"] = {"curve": 0, "L": L, "num_chunks": 1024, "sha256": affine_digest(0, r),
                                              "results": affine_hex(0, r)}
            print("bn254 batched: %.1f s" % (time.time() - t0), flush=True)
        del pts, sc

    if want("bls12_381_2p22"):
        t0 = time.time()
        n = 1 << 22
        pts = O.gen_points(1, SEED, n)
        sc = O.gen_scalars(1, SEED, n)
        r = O.multiexp_cpu(1, pts, sc)
        doc["bls12_381_2p22"] = {"curve": 1, "n": n, "result": affine_hex(1, r)[0],
                                 "oracle_seconds": round(time.time() - t0, 2), "host_threads": O.ncores()}
        print("bls12-381 2^22: %.1f s" % (time.time() - t0), flush=True)
        del pts, sc

    if want("bn254_amt_10x2p21_2048"):
        t0 = time.time()
        lines, L, chunks = 10, 1 << 21, 2048
        pts = O.gen_points(0, SEED, lines * L)
        sc = O.gen_scalars(0, SEED, L)
        r = O.multiple_multiexp(0, pts, sc, chunks)
        doc["bn254_amt_10x2p21_2048"] = {"curve": 0, "lines": lines, "L": L, "num_chunks": chunks,
                                         "sha256": affine_digest(0, r), "first_results": affine_hex(0, r[:8])}
        print("bn254 AMT shape: %.1f s" % (time.time() - t0), flush=True)

    if want("bls12_381_batched_256x4096"):
        t0 = time.time()
        L, chunks = 1 << 20, 256
        pts = O.gen_points(1, SEED, L)
        sc = O.gen_scalars(1, SEED, L)
        r = O.multiple_multiexp(1, pts, sc, chunks)
        doc["bls12_381_batched_256x4096"] = {"curve": 1, "L": L, "num_chunks": chunks, "sha256": affine_digest(1, r),
                                             "first_results": affine_hex(1, r[:8])}
        print("bls12-381 batched: %.1f s" % (time.time() - t0), flush=True)

    for name, curve, k in (("bn254_g2_2p20", 2, 20), ("bls12_381_g2_2p18", 3, 18)):
        if want(name):
            t0 = time.time()
            n = 1 << k
            pts = O.gen_points(curve, SEED, n)
            sc = O.gen_scalars(curve, SEED, n)
            r = O.multiexp_cpu(curve, pts, sc)
            doc[name] = {"curve": curve, "n": n, "result": affine_hex(curve, r)[0],
                         "oracle_seconds": round(time.time() - t0, 2), "host_threads": O.ncores()}
            print("%s: %.1f s" % (name, time.time() - t0), flush=True)

    with open(OUT, "w") as f:
        json.dump(doc, f, indent=0, sort_keys=True)
        f.write("\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
