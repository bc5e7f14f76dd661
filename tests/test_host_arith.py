"""CPU tests of the field / curve code the CUDA kernels are built from.

csrc/{fp,fp29,ec}.cuh compile for the host with an emulated carry flag (csrc/ptx.cuh), so the very
same formulas run here against the oracle, for both field implementations:
  impl 0  FieldSat  saturated 32-bit limbs (BN254, BLS12-381)
  impl 1  FieldU29  lazy 29-bit limbs (BN254) -- built with MSM_CHECK_BOUNDS, which aborts the
          process on any 64-bit column overflow or negative limb in the lazy-reduction scheme.
  impl 2  FieldSatLazy  saturated 32-bit limbs with values in [0, 2p) (the engine's default), and for the
          G2 curve ids FieldExt2Lazy, the quadratic extension over it (csrc/fp2.cuh)
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from util import FQ, assert_same_points

HERE = os.path.dirname(os.path.abspath(__file__))
# (curve, impl); impl 2 = FieldSatLazy, values in [0, 2p); curves 2, 3 = G2: FieldExt2Lazy (Fq2 over it, csrc/fp2.cuh)
IMPLS = [(0, 0), (0, 1), (1, 0), (0, 2), (1, 2), (2, 2), (3, 2)]


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "host_arith", "host_arith.cpp")
    lib = os.path.join(HERE, "host_arith", "libhost_arith.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMSM_CHECK_BOUNDS", "-fPIC", "-shared", "-o", lib, src])
    h = ctypes.CDLL(lib)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    h.host_fq_op.argtypes = [i32, i32, i32, vp, vp, vp, vp, sz]
    h.host_ec_op.argtypes = [i32, i32, i32, vp, vp, vp, sz]
    h.host_madd_chain.argtypes = [i32, i32, vp, sz, sz, sz, vp]
    u32 = ctypes.c_uint32
    h.host_ba_rounds.argtypes = [i32, vp, sz, vp, vp, u32, u32, u32, u32, vp, vp]
    h.host_ba_rounds.restype = ctypes.c_long
    return h


def _rand_fq(oracle, curve, n, rng):
    comps = 2 if curve >= 2 else 1  # Fq2 elements are c0 | c1
    fb = FQ[curve] // comps
    p = int.from_bytes(oracle.constant(curve, 0)[:fb].tobytes(), "little")
    vals = [int.from_bytes(rng.bytes(fb + 8), "little") % p for _ in range(n * comps)]
    # extremes: 0, 1, p-1, all-ones limb patterns below p
    vals[:6] = [0, 1, p - 1, p - 2, (1 << (p.bit_length() - 1)) - 1, (p >> 1)]
    return np.frombuffer(b"".join(v.to_bytes(fb, "little") for v in vals), dtype=np.uint8).copy()


@pytest.mark.parametrize("curve,impl", IMPLS)
def test_field_ops_match_oracle(host, oracle, curve, impl):
    rng = np.random.default_rng(99 + curve)
    n = 3000
    a, b = _rand_fq(oracle, curve, n, rng), _rand_fq(oracle, curve, n, rng)
    r2 = oracle.constant(curve, 2)
    for op in range(9):
        aa = a.copy()
        if op == 7:
            aa[: FQ[curve]] = a[FQ[curve]: 2 * FQ[curve]]
        want = oracle.fq_op(curve, op, aa, b)
        got = np.zeros_like(aa)
        rc = host.host_fq_op(curve, impl, op, aa.ctypes.data, b.ctypes.data, r2.ctypes.data, got.ctypes.data, n)
        assert rc == 0
        bad = np.nonzero((want != got).reshape(n, -1).any(axis=1))[0]
        assert bad.size == 0, f"op {op}: rows {bad[:5]}"


@pytest.mark.parametrize("curve,impl", IMPLS)
def test_curve_ops_match_oracle(host, oracle, curve, impl):
    n, fq = 300, FQ[curve]
    pts = oracle.gen_points(curve, 7, n)
    sc = oracle.gen_scalars(curve, 9, n)
    jac = np.stack([oracle.scalar_mul(curve, pts[i], sc[i]) for i in range(n)])
    jac2 = np.stack([oracle.scalar_mul(curve, pts[(i * 7 + 3) % n], sc[(i + 1) % n]) for i in range(n)])
    jac2[0] = jac[0]      # equal inputs -> doubling branch
    jac2[2] = 0           # infinity operands
    jac[3] = 0
    one = oracle.constant(curve, 1)
    lifted = np.zeros((n, 3 * fq), dtype=np.uint8)
    lifted[:, : 2 * fq] = pts
    lifted[:, 2 * fq:] = one
    neg = pts.copy()
    neg[:, fq:] = oracle.fq_op(curve, 8, pts[:, fq:].copy())

    def run(op, a, b):
        got = np.zeros_like(a)
        rc = host.host_ec_op(curve, impl, op, a.ctypes.data, None if b is None else b.ctypes.data, got.ctypes.data, n)
        assert rc == 0
        return got

    assert_same_points(oracle, curve, run(0, jac, jac2), oracle.ec_op(curve, 0, jac, jac2), "add")
    assert_same_points(oracle, curve, run(1, jac, pts), oracle.ec_op(curve, 1, jac, pts), "madd")
    assert_same_points(oracle, curve, run(6, jac, pts), oracle.ec_op(curve, 1, jac, neg), "madd negated")
    assert_same_points(oracle, curve, run(2, jac, None), oracle.ec_op(curve, 2, jac), "dbl")
    assert_same_points(oracle, curve, run(1, lifted, pts), oracle.ec_op(curve, 2, lifted), "madd P+P")
    assert oracle.to_affine(curve, run(1, lifted, neg))[1].all(), "madd P + (-P)"
    assert oracle.to_affine(curve, run(6, lifted, pts))[1].all(), "madd P - P via sign bit"
    assert_same_points(oracle, curve, run(3, lifted, pts), oracle.ec_op(curve, 2, lifted), "mdbl")
    # to_affine (Montgomery) and the resident-copy round trip
    aff = run(4, jac, None)
    wa, _ = oracle.to_affine(curve, jac, mont_out=True)
    assert (aff[:, : 2 * fq] == wa).all()
    rt = run(7, lifted, pts)
    assert (rt[:, : 2 * fq] == pts).all()
    # small scalar multiples
    ks = (np.arange(n, dtype=np.uint32) * 2654435761 % 70000).astype(np.uint32)
    ks[:4] = [0, 1, 2, 65535]
    got = run(5, jac, ks)
    want = np.stack([oracle.scalar_mul(curve, wa[i] if True else None, np.frombuffer(int(ks[i]).to_bytes(32, "little"), dtype=np.uint8))
                     if jac[i, 2 * fq:].any() else np.zeros(3 * fq, dtype=np.uint8) for i in range(n)])
    assert_same_points(oracle, curve, got, want, "mul_small")


@pytest.mark.parametrize("curve,impl", IMPLS)
def test_long_madd_chains_keep_invariants(host, oracle, curve, impl):
    """Thousands of consecutive mixed additions per accumulator (the bucket loop) with mixed signs,
    repeated points (doubling / cancellation) and identity bases; under MSM_CHECK_BOUNDS this is
    the proof-by-execution of the lazy-reduction invariants in ec.cuh."""
    m, steps, lanes = 97, 1500, 12
    pts = oracle.gen_points(curve, 21, m)
    pts[5] = 0  # identity base
    out = np.zeros((lanes, 3 * FQ[curve]), dtype=np.uint8)
    assert host.host_madd_chain(curve, impl, pts.ctypes.data, m, steps, lanes, out.ctypes.data) == 0
    # recompute with the oracle
    fq = FQ[curve]
    neg = pts.copy()
    neg[:, fq:] = oracle.fq_op(curve, 8, pts[:, fq:].copy())
    want = np.zeros_like(out)
    for i in range(lanes):
        acc = np.zeros((1, 3 * fq), dtype=np.uint8)
        for k in range(steps):
            idx = (i * 7 + k * (i + 1)) % m
            src = neg if ((k ^ i) & 1) else pts
            acc = oracle.ec_op(curve, 1, acc, src[idx: idx + 1].copy())
        want[i] = acc[0]
    assert_same_points(oracle, curve, out, want, "madd chain")


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("rounds,T,m_max", [(1, 32, 5), (3, 64, 64), (5, 64, 2), (9, 32, 1000), (2, 96, 1)])
def test_affine_halving_rounds(host, oracle, curve, rounds, T, m_max):
    """csrc/bucket_affine.cuh, the device function of one thread run thread by thread on the host: after any
    number of halving rounds the points left in a bucket still sum to the bucket's signed entries.  Buckets:
    empty, single, odd and even sizes, one heavy bucket spanning many threads and batches, repeated points
    (P + P: tangent), P and -P neighbours (cancel), identity bases (0,0), an all-identity bucket."""
    fq = FQ[curve]
    rng = np.random.default_rng(1000 * curve + rounds)
    m = 61
    pts = oracle.gen_points(curve, 31, m)
    pts[7] = 0
    pts[8] = 0
    neg = pts.copy()
    neg[:, fq:] = oracle.fq_op(curve, 8, pts[:, fq:].copy())
    NEG = 1 << 31
    buckets = [[], [3], [4, 4], [5, 5 | NEG], [7, 9], [9, 7], [7, 8], [7], [10, 10, 10, 10], [11, 11 | NEG, 12],
               [13, 14, 13 | NEG, 14 | NEG, 15], [], [int(x) for x in rng.integers(0, m, 301)],
               [int(x) | (NEG if s else 0) for x, s in zip(rng.integers(0, m, 40), rng.integers(0, 2, 40))]]
    for _ in range(30):
        k = int(rng.integers(0, 12))
        buckets.append([int(x) | (NEG if s else 0) for x, s in zip(rng.integers(0, m, k), rng.integers(0, 2, k))])
    NB = len(buckets)
    off0 = np.zeros(NB + 1, dtype=np.uint32)
    off0[1:] = np.cumsum([len(b) for b in buckets])
    entries = np.array([e for b in buckets for e in b], dtype=np.uint32)
    out_pts = np.zeros((len(entries) + 1, 2 * fq), dtype=np.uint8)
    out_off = np.zeros(NB + 1, dtype=np.uint32)
    left = host.host_ba_rounds(curve, pts.ctypes.data, m, entries.ctypes.data, off0.ctypes.data, NB, rounds, T, m_max,
                               out_pts.ctypes.data, out_off.ctypes.data)
    assert left >= 0
    n = off0[1:] - off0[:-1]
    for _ in range(rounds):
        n = (n + 1) // 2
    assert (out_off[1:] - out_off[:-1] == n).all() and left == int(n.sum())
    one = oracle.constant(curve, 1)

    def total(rows):
        acc = np.zeros((1, 3 * fq), dtype=np.uint8)
        for r in rows:
            if not r.any():
                continue
            acc = oracle.ec_op(curve, 1, acc, r.reshape(1, -1).copy())
        return acc

    got = np.concatenate([total(out_pts[out_off[g]: out_off[g + 1]]) for g in range(NB)])
    want = np.concatenate([total([(neg if e & NEG else pts)[e & 0x7FFFFFFF] for e in b]) for b in buckets])
    assert_same_points(oracle, curve, got, want, "bucket sums after the halving rounds")
