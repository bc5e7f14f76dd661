"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of liboracle_msm.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_msm.so")
_lib = None

BN254_G1 = 0
BLS12_381_G1 = 1
FQ_BYTES = {0: 32, 1: 48, 2: 64, 3: 96}  # bytes per coordinate (2, 3: G2 over Fq2)


def build(force=False):
    src = os.path.join(_HERE, "msm_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        vp, sz, i32, u32, u64 = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32,
                                 ctypes.c_uint64)
        _lib.oracle_constant.argtypes = [i32, i32, vp]
        _lib.oracle_fq_op.argtypes = [i32, i32, vp, vp, vp, sz]
        _lib.oracle_ec_op.argtypes = [i32, i32, vp, vp, vp, sz]
        _lib.oracle_to_affine.argtypes = [i32, vp, sz, i32, vp, vp]
        _lib.oracle_on_curve.argtypes = [i32, vp, sz]
        _lib.oracle_multiexp_cpu.argtypes = [i32, vp, vp, sz, i32, vp]
        _lib.oracle_msm_naive.argtypes = [i32, vp, vp, sz, vp]
        _lib.oracle_multiple_multiexp.argtypes = [i32, vp, sz, vp, sz, u32, i32, vp]
        _lib.oracle_scalar_mul.argtypes = [i32, vp, vp, vp]
        _lib.oracle_gen_scalars.argtypes = [i32, u64, sz, sz, vp]
        _lib.oracle_gen_points.argtypes = [i32, u64, sz, sz, i32, vp]
        _lib.oracle_ec_fft.argtypes = [i32, vp, u32, vp]
        _lib.oracle_fr_op.argtypes = [i32, i32, vp, vp, vp, sz]
        _lib.oracle_fr_fft.argtypes = [i32, vp, u32, vp]
        _lib.oracle_window_for.argtypes = [sz]
        _lib.oracle_window_for.restype = ctypes.c_uint
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


def ncores():
    return os.cpu_count() or 1


def constant(curve, which):
    size = {0: FQ_BYTES[curve], 1: FQ_BYTES[curve], 2: FQ_BYTES[curve], 3: 8,
            4: 2 * FQ_BYTES[curve], 5: 32, 6: FQ_BYTES[curve]}[which]
    out = np.zeros(size, dtype=np.uint8)
    rc = lib().oracle_constant(curve, which, _ptr(out))
    assert rc == 0
    return out


def fq_op(curve, op, a, b=None):
    """a, b: uint8 arrays [count, FQ_BYTES]; op numbering as in msm_oracle.cpp."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    count = a.size // FQ_BYTES[curve]
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros_like(a)
    rc = lib().oracle_fq_op(curve, op, _ptr(a), _ptr(bb), _ptr(out), count)
    assert rc == 0
    return out


def ec_op(curve, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    count = a.size // (3 * FQ_BYTES[curve])
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros_like(a)
    rc = lib().oracle_ec_op(curve, op, _ptr(a), _ptr(bb), _ptr(out), count)
    assert rc == 0
    return out


def to_affine(curve, jac, mont_out=False):
    """jac: uint8 [count, 3*FQ]; returns (xy uint8 [count, 2*FQ], inf uint8 [count])."""
    jac = np.ascontiguousarray(jac, dtype=np.uint8)
    count = jac.size // (3 * FQ_BYTES[curve])
    out = np.zeros((count, 2 * FQ_BYTES[curve]), dtype=np.uint8)
    inf = np.zeros(count, dtype=np.uint8)
    rc = lib().oracle_to_affine(curve, _ptr(jac), count, 1 if mont_out else 0, _ptr(out), _ptr(inf))
    assert rc == 0
    return out, inf


def on_curve(curve, aff):
    aff = np.ascontiguousarray(aff, dtype=np.uint8)
    count = aff.size // (2 * FQ_BYTES[curve])
    return lib().oracle_on_curve(curve, _ptr(aff), count) == 0


def multiexp_cpu(curve, bases, exps, nthreads=None):
    """The reference's CPU multiexp.  Returns Jacobian bytes; raises on identity base."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    exps = np.ascontiguousarray(exps, dtype=np.uint8)
    n = exps.size // 32
    assert bases.size // (2 * FQ_BYTES[curve]) >= n
    out = np.zeros(3 * FQ_BYTES[curve], dtype=np.uint8)
    rc = lib().oracle_multiexp_cpu(curve, _ptr(bases), _ptr(exps), n, nthreads or ncores(), _ptr(out))
    if rc == -1:
        raise ValueError("Encountered an identity element in the CRS.")
    assert rc == 0, rc
    return out


def msm_naive(curve, bases, exps):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    exps = np.ascontiguousarray(exps, dtype=np.uint8)
    n = exps.size // 32
    out = np.zeros(3 * FQ_BYTES[curve], dtype=np.uint8)
    rc = lib().oracle_msm_naive(curve, _ptr(bases), _ptr(exps), n, _ptr(out))
    assert rc == 0
    return out


def multiple_multiexp(curve, bases, exps, num_chunks, nthreads=None):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    exps = np.ascontiguousarray(exps, dtype=np.uint8)
    L = exps.size // 32
    n_bases = bases.size // (2 * FQ_BYTES[curve])
    num_lines = n_bases // L
    out = np.zeros((num_lines * num_chunks, 3 * FQ_BYTES[curve]), dtype=np.uint8)
    rc = lib().oracle_multiple_multiexp(curve, _ptr(bases), n_bases, _ptr(exps), L, num_chunks,
                                        nthreads or ncores(), _ptr(out))
    assert rc == 0, rc
    return out


def scalar_mul(curve, base_aff, scalar32):
    base_aff = np.ascontiguousarray(base_aff, dtype=np.uint8)
    scalar32 = np.ascontiguousarray(scalar32, dtype=np.uint8)
    out = np.zeros(3 * FQ_BYTES[curve], dtype=np.uint8)
    rc = lib().oracle_scalar_mul(curve, _ptr(base_aff), _ptr(scalar32), _ptr(out))
    assert rc == 0
    return out


def gen_scalars(curve, seed, n, start=0):
    out = np.zeros((n, 32), dtype=np.uint8)
    rc = lib().oracle_gen_scalars(curve, seed, start, n, _ptr(out))
    assert rc == 0
    return out


def gen_points(curve, seed, n, start=0, nthreads=None):
    out = np.zeros((n, 2 * FQ_BYTES[curve]), dtype=np.uint8)
    rc = lib().oracle_gen_points(curve, seed, start, n, nthreads or ncores(), _ptr(out))
    assert rc == 0
    return out


def fr_op(curve, op, a, b=None):
    """Scalar field: op 0 to Montgomery, 1 from Montgomery, 2 Montgomery product.  [count, 32] uint8."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros_like(a)
    rc = lib().oracle_fr_op(curve, op, _ptr(a), _ptr(bb), _ptr(out), a.size // 32)
    assert rc == 0
    return out


def ec_fft(curve, jac, omega_mont):
    """serial_ec_fft restated: returns the transformed copy of jac ([n, 3*FQ] uint8, n a power of two)."""
    out = np.ascontiguousarray(jac, dtype=np.uint8).copy()
    n = out.size // (3 * FQ_BYTES[curve])
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    om = np.ascontiguousarray(omega_mont, dtype=np.uint8)
    rc = lib().oracle_ec_fft(curve, _ptr(out), log_n, _ptr(om))
    assert rc == 0
    return out


def fr_fft(curve, elems_mont, omega_mont):
    """serial_fft restated: returns the transformed copy of elems_mont ([n, 32] uint8 Fr Montgomery, n = 2^k)."""
    out = np.ascontiguousarray(elems_mont, dtype=np.uint8).copy()
    n = out.size // 32
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    om = np.ascontiguousarray(omega_mont, dtype=np.uint8)
    rc = lib().oracle_fr_fft(curve, _ptr(out), log_n, _ptr(om))
    assert rc == 0
    return out


def window_for(n):
    return lib().oracle_window_for(n)
