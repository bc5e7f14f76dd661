"""ctypes binding of include/msm_b200.h.

There is no fallback of any kind here: if libmsm_b200.so is missing or does not export a symbol
the header declares, importing callers get an exception.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = os.environ.get("MSM_B200_LIB") or os.path.join(_HERE, "libmsm_b200.so")  # override: A/B builds only
_HEADER = os.path.join(_ROOT, "include", "msm_b200.h")

BN254_G1 = 0
BLS12_381_G1 = 1
BN254_G2 = 2       # engine extension: the same API over Fq2 coordinates
BLS12_381_G2 = 3

MSM_OK, MSM_ERR_INVALID, MSM_ERR_CUDA, MSM_ERR_BUSY, MSM_ERR_ABORTED, MSM_ERR_NO_DEVICE, MSM_ERR_TOO_LARGE = range(7)


def fq_bytes(curve: int) -> int:
    return {BN254_G1: 32, BLS12_381_G1: 48, BN254_G2: 64, BLS12_381_G2: 96}[curve]  # bytes per coordinate


class EcError(Exception):
    """ec_gpu_program::EcError (ec-gpu-program/src/lib.rs:11-32)."""


class EcErrorSimple(EcError):
    pass


class EcErrorAborted(EcError):
    pass


class EcErrorGpuTools(EcError):
    pass


class CudaError(Exception):
    """rustacuda::error::CudaError as surfaced by ag_cuda_ec (CudaResult)."""

    def __init__(self, name, detail=""):
        super().__init__(f"{name}: {detail}" if detail else name)
        self.name = name


class Timings(ctypes.Structure):
    _fields_ = [
        ("h2d_ms", ctypes.c_float),
        ("sort_ms", ctypes.c_float),
        ("accumulate_ms", ctypes.c_float),
        ("reduce_ms", ctypes.c_float),
        ("total_ms", ctypes.c_float),
        ("window_bits", ctypes.c_uint32),
        ("num_windows", ctypes.c_uint32),
        ("num_entries", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("scatter_passes", ctypes.c_uint32),
        ("sub_batches", ctypes.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PlanInfo(ctypes.Structure):
    """msm_plan_info (include/msm_b200.h)."""
    _fields_ = [
        ("window_bits", ctypes.c_uint32),
        ("num_windows", ctypes.c_uint32),
        ("buckets", ctypes.c_uint32),
        ("sub_batches", ctypes.c_uint32),
        ("by_task", ctypes.c_uint32),
        ("sub_first", ctypes.c_uint32 * 9),
        ("slice_len", ctypes.c_uint32),
        ("slices", ctypes.c_uint32),
        ("wave_slices", ctypes.c_uint32),
        ("waves", ctypes.c_uint32),
        ("sort_mode", ctypes.c_uint32),
        ("reduce_q", ctypes.c_uint32),
        ("digits_max", ctypes.c_uint64),
        ("scratch_bytes", ctypes.c_uint64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["sub_first"] = list(self.sub_first)
        return d


def describe_plan(curve: int, n_scalars: int, n_lines: int = 1, num_chunks: int = 1, table_window_bits: int = 0,
                  sub_batches: int = 1, growth: float = 2.0) -> dict:
    """msm_plan_describe: the launch plan the engine makes for a call of this shape (host arithmetic, no GPU needed)."""
    info = PlanInfo()
    rc = load_library().msm_plan_describe(curve, n_scalars, n_lines, num_chunks, table_window_bits, sub_batches,
                                          float(growth), ctypes.byref(info))
    if rc != MSM_OK:
        raise CudaError("InvalidValue" if rc == MSM_ERR_INVALID else f"msm_plan_describe: error {rc}")
    return info.as_dict()


def pipeline_shape(n_scalars: int, n_lines: int = 1, num_chunks: int = 1, h2d_gbs: float = 0.0, device_ms: float = 0.0):
    """msm_pipeline_shape: (sub-batches, growth factor) of the pipelined scalar upload for a call of this shape."""
    n, g = ctypes.c_uint32(), ctypes.c_double()
    rc = load_library().msm_pipeline_shape(n_scalars, n_lines, num_chunks, h2d_gbs, device_ms, ctypes.byref(n), ctypes.byref(g))
    if rc != MSM_OK:
        raise CudaError("InvalidValue")
    return n.value, g.value


def library_path() -> str:
    return _LIB


def build_library(force: bool = False) -> str:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s"]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return _LIB


def header_symbols() -> list[str]:
    """Every function the C header declares."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msm_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise ImportError(
            f"{_LIB} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or __graft_entry__.build()). There is no CPU fallback."
        )
    lib = ctypes.CDLL(_LIB)
    vp, sz, i32, u32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
    pp = ctypes.POINTER(ctypes.c_void_p)
    sig = {
        "msm_device_count": ([], i32),
        "msm_ctx_create": ([i32, ctypes.POINTER(i32), i32, pp], i32),
        "msm_ctx_destroy": ([vp], i32),
        "msm_ctx_num_devices": ([vp], i32),
        "msm_set_abort_flag": ([vp, vp], i32),
        "msm_last_error": ([vp], ctypes.c_char_p),
        "msm_last_timings": ([vp, ctypes.POINTER(Timings)], i32),
        "msm_set_window_bits": ([vp, u32], i32),
        "msm_plan_describe": ([i32, sz, u32, u32, u32, u32, ctypes.c_double, ctypes.POINTER(PlanInfo)], i32),
        "msm_pipeline_shape": ([sz, u32, u32, ctypes.c_float, ctypes.c_float, ctypes.POINTER(u32), ctypes.POINTER(ctypes.c_double)], i32),
        "msm_field_impl": ([vp], ctypes.c_char_p),
        "msm_bases_upload": ([vp, vp, sz, pp], i32),
        "msm_bases_upload_sharded": ([vp, vp, sz, pp], i32),
        "msm_bases_from_device": ([vp, vp, sz, pp], i32),
        "msm_bases_precompute": ([vp, vp, u32], i32),
        "msm_bases_precompute_chunked": ([vp, vp, sz], i32),
        "msm_bases_table_window": ([vp], u32),
        "msm_bases_set_table_policy": ([vp, vp, i32], i32),
        "msm_bases_size_bytes": ([vp], sz),
        "msm_bases_num_points": ([vp], sz),
        "msm_bases_free": ([vp], i32),
        "msm_multiple_multiexp": ([vp, vp, vp, sz, u32, u32, i32, vp], i32),
        "msm_multiple_multiexp_device": ([vp, vp, vp, sz, u32, vp], i32),
        "msm_multiple_multiexp_device_timed": ([vp, vp, vp, sz, u32, vp, u32, ctypes.POINTER(ctypes.c_float),
                                                ctypes.POINTER(ctypes.c_float)], i32),
        "msm_set_stream": ([vp, vp], i32),
        "msm_scalars_from_montgomery_device": ([vp, vp, sz, vp], i32),
        "msm_multiple_multiexp_montgomery": ([vp, vp, vp, sz, u32, vp], i32),
        "msm_multiexp": ([vp, vp, vp, sz, vp], i32),
        "msm_multiexp_resident": ([vp, vp, sz, vp, sz, vp], i32),
        "msm_sum_points_device": ([vp, vp, sz, vp], i32),
        "msm_ec_fft": ([vp, vp, u32, vp, u32], i32),
        "msm_ec_fft_device": ([vp, vp, u32, vp, u32], i32),
        "msm_scalar_fft": ([vp, vp, u32, vp], i32),
        "msm_scalar_fft_device": ([vp, vp, u32, vp], i32),
        "msm_to_affine": ([vp, vp, sz, i32, vp, vp], i32),
        "msm_synth_points_device": ([vp, u64, sz, sz, vp], i32),
        "msm_synth_scalars_device": ([vp, u64, sz, sz, vp], i32),
        "msm_test_fq_op": ([vp, i32, vp, vp, vp, sz], i32),
        "msm_test_ec_op": ([vp, i32, vp, vp, vp, sz], i32),
        "msm_device_alloc": ([vp, sz, pp], i32),
        "msm_device_free": ([vp, vp], i32),
        "msm_memcpy_h2d": ([vp, vp, vp, sz], i32),
        "msm_memcpy_d2h": ([vp, vp, vp, sz], i32),
        "msm_host_register": ([vp, sz], i32),
        "msm_host_unregister": ([vp], i32),
        "msm_version": ([], ctypes.c_char_p),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc: int, ctx=None, cuda_style: bool = False):
    """Map an msm_status to the reference's error types."""
    if rc == MSM_OK:
        return
    lib = load_library()
    msg = lib.msm_last_error(ctx)
    msg = msg.decode() if msg else ""
    if cuda_style:
        # ag_cuda_ec returns CudaResult<_>
        if rc == MSM_ERR_BUSY:
            raise CudaError("ContextAlreadyInUse", msg)
        if rc == MSM_ERR_NO_DEVICE:
            raise CudaError("NoDevice", msg or "No working GPUs found!")
        if rc == MSM_ERR_INVALID:
            raise CudaError("InvalidValue", msg)
        if rc == MSM_ERR_TOO_LARGE:
            raise CudaError("InvalidValue", "size exceeds the u32 index space")
        raise CudaError("UnknownError", msg)
    if rc == MSM_ERR_ABORTED:
        raise EcErrorAborted("GPU call was aborted!")
    if rc == MSM_ERR_NO_DEVICE:
        raise EcErrorSimple("No working GPUs found!")
    if rc == MSM_ERR_CUDA:
        raise EcErrorGpuTools(msg)
    if rc == MSM_ERR_BUSY:
        raise EcErrorGpuTools("context already in use")
    raise EcErrorSimple(msg or f"msm_status {rc}")
