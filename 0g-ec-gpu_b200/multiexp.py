"""Mirror of ag_cuda_ec::multiexp (ag-cuda-ec/src/multiexp.rs) over the C ABI.

Same names, argument meaning and error behaviour as the Rust API; arrays are numpy uint8 views of
exactly the reference's memory layouts:

  bases      [n, 2*FQ]  {x, y} Montgomery little-endian  (GpuRepr, ag-types/src/impls.rs:48-58)
  exponents  [L, 32]    canonical little-endian BigInt<4> (PrimeFieldRepr::to_bigint)
  result     [lines*chunks, 3*FQ]  Jacobian {x, y, z} Montgomery (Vec<Curve>)
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib
from ._lib import BN254_G1, check, fq_bytes, load_library


class Workspace:
    """CudaWorkspace (ag-cuda-proxy/src/module.rs:13-62): one engine context."""

    def __init__(self, curve: int = BN254_G1, devices=None):
        lib = load_library()
        self.curve = curve
        self._h = ctypes.c_void_p()
        if devices is None:
            ids, n = None, 1  # the reference pins device 0 (module.rs:27)
        else:
            devices = list(devices)
            ids, n = (ctypes.c_int * len(devices))(*devices), len(devices)
        rc = lib.msm_ctx_create(curve, ids, n, ctypes.byref(self._h))
        check(rc, None, cuda_style=True)

    @property
    def handle(self):
        return self._h

    def num_devices(self) -> int:
        return load_library().msm_ctx_num_devices(self._h)

    def timings(self) -> dict:
        t = _lib.Timings()
        load_library().msm_last_timings(self._h, ctypes.byref(t))
        return t.as_dict()

    def set_window_bits(self, c: int):
        check(load_library().msm_set_window_bits(self._h, c), self._h, cuda_style=True)

    def close(self):
        if self._h:
            load_library().msm_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceData:
    """DeviceData (ag-cuda-proxy/src/params.rs:173-218): resident bases, freed on drop."""

    def __init__(self, workspace: Workspace, handle):
        self.workspace = workspace
        self._h = handle

    @property
    def _as_parameter_(self):  # lets ctypes take a DeviceData wherever the C ABI wants an msm_bases*
        return self._h

    def set_table_policy(self, policy: int):
        """Engine extension (msm_bases_set_table_policy): 0 off, 1 lazy (default), 2 eager."""
        check(load_library().msm_bases_set_table_policy(self.workspace.handle, self._h, policy), self.workspace.handle,
              cuda_style=True)

    def table_window(self) -> int:
        return load_library().msm_bases_table_window(self._h)

    def size(self) -> int:
        return load_library().msm_bases_size_bytes(self._h)

    def num_points(self) -> int:
        return load_library().msm_bases_num_points(self._h)

    def precompute(self, window_bits: int = 0) -> int:
        """Engine extension (msm_bases_precompute): build the window table for repeated large
        single MSMs over these bases.  Returns the table's window size."""
        lib = load_library()
        check(lib.msm_bases_precompute(self.workspace.handle, self._h, window_bits), self.workspace.handle,
              cuda_style=True)
        return lib.msm_bases_table_window(self._h)

    def precompute_chunked(self, chunk_len: int) -> int:
        """Engine extension (msm_bases_precompute_chunked): window table sized for
        multiple_multiexp calls whose tasks have chunk_len points each (the per-segment commitment
        and AMT shapes).  Returns the table's window size."""
        lib = load_library()
        check(lib.msm_bases_precompute_chunked(self.workspace.handle, self._h, chunk_len), self.workspace.handle,
              cuda_style=True)
        return lib.msm_bases_table_window(self._h)

    def free(self):
        if self._h:
            load_library().msm_bases_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# construct_workspace! (ag-cuda-workspace-macro/src/lib.rs:58-78): one GLOBAL, one per-thread LOCAL
_GLOBAL: Workspace | None = None
_GLOBAL_LOCK = threading.Lock()
_LOCAL = threading.local()
_DEFAULT_CURVE = BN254_G1  # Cargo feature `bn254` is the default (ag-cuda-ec/Cargo.toml:35-38)


def init_global_workspace(curve: int | None = None) -> Workspace:
    global _GLOBAL
    with _GLOBAL_LOCK:
        want = _DEFAULT_CURVE if curve is None else curve
        if _GLOBAL is None or _GLOBAL.curve != want:
            _GLOBAL = Workspace(want)
        return _GLOBAL


def init_local_workspace(curve: int | None = None) -> Workspace:
    want = _DEFAULT_CURVE if curve is None else curve
    ws = getattr(_LOCAL, "ws", None)
    if ws is None or ws.curve != want:
        ws = Workspace(want)
        _LOCAL.ws = ws
    return ws


def _as_u8(a, row_bytes, what):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        a = a.view(np.uint8)
    a = a.reshape(-1)
    if a.size % row_bytes:
        raise ValueError(f"{what}: byte length {a.size} is not a multiple of {row_bytes}")
    return a


def upload_multiexp_bases(workspace: Workspace, bases) -> DeviceData:
    """ag-cuda-ec/src/multiexp.rs:12-19."""
    pt = 2 * fq_bytes(workspace.curve)
    b = _as_u8(bases, pt, "bases")
    h = ctypes.c_void_p()
    rc = load_library().msm_bases_upload(workspace.handle, b.ctypes.data, b.size // pt, ctypes.byref(h))
    check(rc, workspace.handle, cuda_style=True)
    return DeviceData(workspace, h)


def multiple_multiexp(workspace: Workspace, bases_gpu: DeviceData, exponents, num_chunks: int,
                      window_size: int, neg_is_cheap: bool) -> np.ndarray:
    """ag-cuda-ec/src/multiexp.rs:22-81.  Returns [num_lines*num_chunks, 3*FQ] uint8."""
    e = _as_u8(exponents, 32, "exponents")
    L = e.size // 32
    if L == 0:
        raise ZeroDivisionError("attempt to divide by zero")  # num_bases / exponents.len() panics
    num_lines = bases_gpu.num_points() // L
    out = np.zeros((num_lines * num_chunks, 3 * fq_bytes(workspace.curve)), dtype=np.uint8)
    rc = load_library().msm_multiple_multiexp(workspace.handle, bases_gpu._h, e.ctypes.data, L, num_chunks,
                                              window_size, 1 if neg_is_cheap else 0, out.ctypes.data)
    check(rc, workspace.handle, cuda_style=True)
    return out


def multiple_multiexp_montgomery(workspace: Workspace, bases_gpu: DeviceData, exponents_mont, num_chunks: int) -> np.ndarray:
    """Engine extension: `exponents_mont` are Fr elements still in
    Montgomery form ([L, 32] uint8, arkworks' in-memory layout); the conversion that
    PrimeFieldRepr::to_bigint does on the host runs on the device instead."""
    e = _as_u8(exponents_mont, 32, "exponents")
    L = e.size // 32
    num_lines = bases_gpu.num_points() // L
    out = np.zeros((num_lines * num_chunks, 3 * fq_bytes(workspace.curve)), dtype=np.uint8)
    rc = load_library().msm_multiple_multiexp_montgomery(workspace.handle, bases_gpu._h, e.ctypes.data, L, num_chunks,
                                                         out.ctypes.data)
    check(rc, workspace.handle, cuda_style=True)
    return out


# #[auto_workspace] (ag-cuda-workspace-macro/src/lib.rs:8-55): f_st uses GLOBAL, f_mt uses LOCAL
def upload_multiexp_bases_st(bases, curve: int | None = None) -> DeviceData:
    return upload_multiexp_bases(init_global_workspace(curve), bases)


def upload_multiexp_bases_mt(bases, curve: int | None = None) -> DeviceData:
    return upload_multiexp_bases(init_local_workspace(curve), bases)


def multiple_multiexp_st(bases_gpu: DeviceData, exponents, num_chunks, window_size, neg_is_cheap):
    return multiple_multiexp(bases_gpu.workspace, bases_gpu, exponents, num_chunks, window_size, neg_is_cheap)


def multiple_multiexp_mt(bases_gpu: DeviceData, exponents, num_chunks, window_size, neg_is_cheap):
    return multiple_multiexp(bases_gpu.workspace, bases_gpu, exponents, num_chunks, window_size, neg_is_cheap)
