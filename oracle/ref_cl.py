"""oracle/ref_cl.py -- TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libref_<curve>.so,
the reference's own device sources compiled for the host by oracle/build_ref.py."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build_ref

_NAMES = {0: "bn254", 1: "bls12_381"}
_FQ = {0: 32, 1: 48}
_libs = {}


def available() -> bool:
    return os.path.isdir(build_ref.CL) or all(
        os.path.exists(os.path.join(build_ref.OUT, "libref_%s.so" % n)) for n in _NAMES.values())


def lib(curve: int):
    if curve not in _libs:
        built = build_ref.build()
        path = (built or {}).get(_NAMES[curve]) or os.path.join(build_ref.OUT, "libref_%s.so" % _NAMES[curve])
        l = ctypes.CDLL(path)
        vp, sz, i32, u32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint
        l.ref_fq_op.argtypes = [i32, vp, vp, vp, sz]
        l.ref_ec_op.argtypes = [i32, vp, vp, vp, sz]
        l.ref_multiple_multiexp.argtypes = [vp, sz, vp, sz, u32, u32, i32, vp]
        l.ref_fr_fft.argtypes = [vp, vp, u32]
        l.ref_fq2_op.argtypes = [i32, vp, vp, vp, sz]
        l.ref_g2_ec_op.argtypes = [i32, vp, vp, vp, sz]
        l.ref_g2_multiple_multiexp.argtypes = [vp, sz, vp, sz, u32, u32, i32, vp]
        l.ref_ec_fft.argtypes = [vp, vp, u32]
        _libs[curve] = l
    return _libs[curve]


def fq_op(curve, op, a, b=None):
    """op: 0 add 1 sub 2 mul 3 sqr 4 double 5 mont 6 unmont (FIELD_* of ag-build/cl/field.cl)."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    bb = a if b is None else np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros_like(a)
    rc = lib(curve).ref_fq_op(op, a.ctypes.data, bb.ctypes.data, out.ctypes.data, a.size // _FQ[curve])
    assert rc == 0
    return out


def ec_op(curve, op, a, b=None):
    """op: 0 POINT_add 1 POINT_add_mixed 2 POINT_double (ag-build/cl/ec.cl)."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.zeros_like(a)
    bp = None if b is None else np.ascontiguousarray(b, dtype=np.uint8).ctypes.data
    rc = lib(curve).ref_ec_op(op, a.ctypes.data, bp, out.ctypes.data, a.size // (3 * _FQ[curve]))
    assert rc == 0
    return out


def multiple_multiexp(curve, bases, exps, num_chunks, window_size, neg_is_cheap):
    """The reference's POINT_multiexp kernel driven with the geometry of
    ag_cuda_ec::multiple_multiexp (ag-cuda-ec/src/multiexp.rs:27-72)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    exps = np.ascontiguousarray(exps, dtype=np.uint8)
    L = exps.size // 32
    n_bases = bases.size // (2 * _FQ[curve])
    out = np.zeros(((n_bases // L) * num_chunks, 3 * _FQ[curve]), dtype=np.uint8)
    rc = lib(curve).ref_multiple_multiexp(bases.ctypes.data, n_bases, exps.ctypes.data, L, num_chunks,
                                          window_size, 1 if neg_is_cheap else 0, out.ctypes.data)
    assert rc == 0
    return out


def fr_fft(curve, elems_mont, omega_mont):
    """The reference's FIELD_radix_fft kernel (ag-build/cl/fft.cl:4-66) under the pass loop of
    SingleFftKernel::radix_fft (ec-gpu-proxy/src/fft.rs:50-136).  Returns the transformed copy."""
    out = np.ascontiguousarray(elems_mont, dtype=np.uint8).copy()
    n = out.size // 32
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    om = np.ascontiguousarray(omega_mont, dtype=np.uint8)
    assert lib(curve).ref_fr_fft(out.ctypes.data, om.ctypes.data, log_n) == 0
    return out


def ec_fft(curve, jac, omegas_mont):
    """The reference's POINT_radix_fft kernel (ag-build/cl/ec-fft.cl:4-76) under the pass loop of
    ag_cuda_ec::ec_fft::radix_ec_fft (ag-cuda-ec/src/ec_fft.rs:13-99).  omegas_mont: [32, 32]."""
    out = np.ascontiguousarray(jac, dtype=np.uint8).copy()
    n = out.size // (3 * _FQ[curve])
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    om = np.ascontiguousarray(omegas_mont, dtype=np.uint8)
    assert om.size == 32 * 32
    assert lib(curve).ref_ec_fft(out.ctypes.data, om.ctypes.data, log_n) == 0
    return out


# ---- G2: the reference's field2.cl / ec.cl / multiexp.cl instantiated over Fq2 the way its SourceBuilder
# would (ag-build/src/source/synthesis.rs:100-110); `curve` is the G1 id (0 / 1) of the same pairing suite.
def fq2_op(curve, op, a, b=None):
    """op: 0 add 1 sub 2 mul 3 sqr 4 double (FIELD2_* of ag-build/cl/field2.cl)."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    bb = a if b is None else np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros_like(a)
    assert lib(curve).ref_fq2_op(op, a.ctypes.data, bb.ctypes.data, out.ctypes.data, a.size // (2 * _FQ[curve])) == 0
    return out


def g2_ec_op(curve, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.zeros_like(a)
    bp = None if b is None else np.ascontiguousarray(b, dtype=np.uint8).ctypes.data
    assert lib(curve).ref_g2_ec_op(op, a.ctypes.data, bp, out.ctypes.data, a.size // (6 * _FQ[curve])) == 0
    return out


def g2_multiple_multiexp(curve, bases, exps, num_chunks, window_size, neg_is_cheap):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    exps = np.ascontiguousarray(exps, dtype=np.uint8)
    L = exps.size // 32
    n_bases = bases.size // (4 * _FQ[curve])
    out = np.zeros(((n_bases // L) * num_chunks, 6 * _FQ[curve]), dtype=np.uint8)
    rc = lib(curve).ref_g2_multiple_multiexp(bases.ctypes.data, n_bases, exps.ctypes.data, L, num_chunks,
                                             window_size, 1 if neg_is_cheap else 0, out.ctypes.data)
    assert rc == 0
    return out
