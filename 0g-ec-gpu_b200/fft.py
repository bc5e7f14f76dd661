"""Mirror of ec_gpu_proxy::fft::{SingleFftKernel, FftKernel} (ec-gpu-proxy/src/fft.rs:19-260) and of
ec_gpu_proxy::ec_fft::SingleEcFftKernel (ec-gpu-proxy/src/ec_fft.rs:19-160) over the C ABI.

  input  [n, 32] uint8  elements of the scalar field Fr in arkworks' in-memory layout (Montgomery,
                        little-endian), n = 2^log_n; transformed in place
  omega  [32] uint8     a primitive n-th root of unity in the same layout
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import BN254_G1, EcErrorAborted, check, fq_bytes, load_library
from .multiexp import Workspace, _as_u8

LOG2_MAX_ELEMENTS = 32  # ec-gpu-proxy/src/fft.rs:14


def _inplace_u8(a, what):
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]):
        raise TypeError(f"{what} must be a C-contiguous uint8 array (it is transformed in place)")
    return a


class FftKernel:
    """FftKernel<F> (fft.rs:139-260): `radix_fft` on the first device, `radix_fft_many` over a list."""

    def __init__(self, workspace: Workspace, maybe_abort=None):
        self._ws = workspace
        self._maybe_abort = maybe_abort

    @classmethod
    def create(cls, devices=None, curve: int = BN254_G1):
        return cls.create_with_abort(devices, None, curve)

    @classmethod
    def create_with_abort(cls, devices, maybe_abort, curve: int = BN254_G1):
        lib = load_library()
        if devices is None:
            devices = list(range(max(lib.msm_device_count(), 0)))
        h = ctypes.c_void_p()
        ids = (ctypes.c_int * len(devices))(*devices) if devices else None
        rc = lib.msm_ctx_create(curve, ids, len(devices), ctypes.byref(h)) if devices else 5
        check(rc, None)  # EcError::Simple("No working GPUs found!"), fft.rs:184-186
        ws = Workspace.__new__(Workspace)
        ws.curve = curve
        ws._h = h
        return cls(ws, maybe_abort)

    @property
    def workspace(self) -> Workspace:
        return self._ws

    def radix_fft(self, input: np.ndarray, omega, log_n: int) -> None:  # noqa: A002
        """fft.rs:200-204 -> SingleFftKernel::radix_fft (fft.rs:50-136)."""
        a = _inplace_u8(input, "input")
        if a.size != 32 << log_n:
            raise ValueError(f"input holds {a.size // 32} elements, log_n = {log_n}")
        if self._maybe_abort is not None and self._maybe_abort():
            raise EcErrorAborted("GPU call was aborted!")
        om = _as_u8(omega, 32, "omega")
        check(load_library().msm_scalar_fft(self._ws.handle, a.ctypes.data, log_n, om.ctypes.data), self._ws.handle)

    def radix_fft_many(self, inputs, omegas, log_ns) -> None:
        """fft.rs:211-259.  The reference spreads the list over its devices; one device serves it here."""
        for a, om, log_n in zip(inputs, omegas, log_ns):
            self.radix_fft(a, om, log_n)


class EcFftKernel:
    """SingleEcFftKernel / EcFftKernel (ec-gpu-proxy/src/ec_fft.rs:19-160): the legacy entry takes
    omega itself and derives omegas[i] = omega^(2^i) (ec_fft.rs:88-93) -- here with the scalar-field
    arithmetic of the engine's own test kernels, not on the host."""

    def __init__(self, workspace: Workspace, maybe_abort=None):
        self._ws = workspace
        self._maybe_abort = maybe_abort

    @classmethod
    def create(cls, devices=None, curve: int = BN254_G1, maybe_abort=None):
        k = FftKernel.create_with_abort(devices, maybe_abort, curve)
        return cls(k.workspace, maybe_abort)

    def radix_ec_fft(self, input: np.ndarray, omegas, log_n: int) -> None:  # noqa: A002
        """`omegas`: [>= log_n, 32] with omegas[i] = omega^(2^i) (what ec_fft.rs:88-93 computes from omega)."""
        a = _inplace_u8(input, "input")
        if a.size != (3 * fq_bytes(self._ws.curve)) << log_n:
            raise ValueError("input size does not match log_n")
        if self._maybe_abort is not None and self._maybe_abort():
            raise EcErrorAborted("GPU call was aborted!")
        om = _as_u8(omegas, 32, "omegas")
        check(load_library().msm_ec_fft(self._ws.handle, a.ctypes.data, log_n, om.ctypes.data, om.size // 32),
              self._ws.handle)
