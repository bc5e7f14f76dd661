"""Mirror of ec_gpu_proxy::multiexp::MultiexpKernel (ec-gpu-proxy/src/multiexp.rs:256-403)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._lib import BN254_G1, EcErrorAborted, check, fq_bytes, load_library
from .multiexp import DeviceData, Workspace, _as_u8


class Worker:
    """threadpool::Worker (ec-gpu-proxy/src/threadpool.rs:19-113).  The engine drives one host
    thread per GPU internally, so the pool only carries the EC_GPU_NUM_THREADS convention."""

    def __init__(self):
        n = os.environ.get("EC_GPU_NUM_THREADS")
        self.num_threads = int(n) if n else (os.cpu_count() or 1)

    def log_num_threads(self) -> int:
        return max(self.num_threads, 1).bit_length() - 1


class MultiexpKernel:
    """One kernel object per device; `multiexp` splits the terms over the devices in contiguous
    chunks of ceil(n / devices) and returns the sum of the per-device results."""

    def __init__(self, workspace: Workspace, maybe_abort=None):
        self._ws = workspace
        self._maybe_abort = maybe_abort
        self._resident = None  # (id(bases), DeviceData-like handle)

    @classmethod
    def create(cls, devices=None, curve: int = BN254_G1):
        """MultiexpKernel::create (multiexp.rs:266-271).  devices: list of CUDA ordinals (None = all)."""
        return cls._create_optional_abort(devices, curve, None)

    @classmethod
    def create_with_abort(cls, devices, maybe_abort, curve: int = BN254_G1):
        """MultiexpKernel::create_with_abort (multiexp.rs:278-283)."""
        return cls._create_optional_abort(devices, curve, maybe_abort)

    @classmethod
    def _create_optional_abort(cls, devices, curve, maybe_abort):
        lib = load_library()
        if devices is None:
            devices = list(range(max(lib.msm_device_count(), 0)))
        h = ctypes.c_void_p()
        ids = (ctypes.c_int * len(devices))(*devices) if devices else None
        rc = lib.msm_ctx_create(curve, ids, len(devices), ctypes.byref(h)) if devices else 5
        check(rc, None)  # EcError::Simple("No working GPUs found!")
        ws = Workspace.__new__(Workspace)
        ws.curve = curve
        ws._h = h
        return cls(ws, maybe_abort)

    def num_kernels(self) -> int:
        return self._ws.num_devices()

    @property
    def workspace(self) -> Workspace:
        return self._ws

    def multiexp(self, pool, bases, exps, skip: int = 0) -> np.ndarray:
        """MultiexpKernel::multiexp (multiexp.rs:372-400): uses bases[skip .. skip + len(exps)).
        Returns one Jacobian point [3*FQ] uint8."""
        del pool
        curve = self._ws.curve
        pt = 2 * fq_bytes(curve)
        e = _as_u8(exps, 32, "exps")
        n = e.size // 32
        b = _as_u8(bases, pt, "bases")
        if skip + n > b.size // pt:
            raise IndexError("range end index out of range for bases")  # slice panic in Rust
        if self._maybe_abort is not None and self._maybe_abort():
            raise EcErrorAborted("GPU call was aborted!")
        out = np.zeros(3 * fq_bytes(curve), dtype=np.uint8)
        rc = load_library().msm_multiexp(self._ws.handle, b.ctypes.data + skip * pt, e.ctypes.data, n,
                                         out.ctypes.data)
        check(rc, self._ws.handle)
        return out

    def upload_bases(self, bases):
        """Engine extension: keep the bases resident, sharded over the devices the way
        parallel_multiexp splits them, so repeated calls skip the per-call upload."""
        curve = self._ws.curve
        pt = 2 * fq_bytes(curve)
        b = _as_u8(bases, pt, "bases")
        h = ctypes.c_void_p()
        check(load_library().msm_bases_upload_sharded(self._ws.handle, b.ctypes.data, b.size // pt, ctypes.byref(h)),
              self._ws.handle)
        return DeviceData(self._ws, h)  # freed on drop, like every other resident handle

    def multiexp_resident(self, resident, exps, skip: int = 0) -> np.ndarray:
        e = _as_u8(exps, 32, "exps")
        out = np.zeros(3 * fq_bytes(self._ws.curve), dtype=np.uint8)
        if self._maybe_abort is not None and self._maybe_abort():
            raise EcErrorAborted("GPU call was aborted!")
        rc = load_library().msm_multiexp_resident(self._ws.handle, resident, skip, e.ctypes.data, e.size // 32,
                                                  out.ctypes.data)
        check(rc, self._ws.handle)
        return out
