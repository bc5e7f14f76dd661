"""Mirror of ag_cuda_ec::ec_fft (ag-cuda-ec/src/ec_fft.rs) and of ec_gpu_proxy's SingleEcFftKernel
(ec-gpu-proxy/src/ec_fft.rs:19-160) over the C ABI.

  input   [n, 3*FQ] uint8   Jacobian {x, y, z} Montgomery (Vec<Curve>), n a power of two; transformed in place
  omegas  [>= log2 n, 32]   omegas[i] = omega^(2^i), arkworks' in-memory Fr (Montgomery, little-endian)
"""
from __future__ import annotations

import numpy as np

from ._lib import check, fq_bytes, load_library
from .multiexp import Workspace, _as_u8, init_global_workspace, init_local_workspace


def radix_ec_fft(workspace: Workspace, input: np.ndarray, omegas) -> None:  # noqa: A002 (the reference's name)
    """ag-cuda-ec/src/ec_fft.rs:13-99: in-place DFT over G1, out[k] = sum_j omega^(j k) in[j]."""
    pt = 3 * fq_bytes(workspace.curve)
    if not (isinstance(input, np.ndarray) and input.dtype == np.uint8 and input.flags["C_CONTIGUOUS"]):
        raise TypeError("input must be a C-contiguous uint8 array (it is transformed in place)")
    if input.size % pt:
        raise ValueError(f"input: byte length {input.size} is not a multiple of {pt}")
    n = input.size // pt
    if n == 0:
        raise ValueError("attempt to calculate the logarithm of zero")  # n.ilog2() panics in the reference
    log_n = n.bit_length() - 1
    assert n == 1 << log_n  # assert_eq!(n, 1 << log_n), ag-cuda-ec/src/ec_fft.rs:21
    om = _as_u8(omegas, 32, "omegas")
    rc = load_library().msm_ec_fft(workspace.handle, input.ctypes.data, log_n, om.ctypes.data, om.size // 32)
    check(rc, workspace.handle, cuda_style=True)


def radix_ec_fft_st(input: np.ndarray, omegas, curve: int | None = None) -> None:  # noqa: A002
    radix_ec_fft(init_global_workspace(curve), input, omegas)


def radix_ec_fft_mt(input: np.ndarray, omegas, curve: int | None = None) -> None:  # noqa: A002
    radix_ec_fft(init_local_workspace(curve), input, omegas)
