"""0g-ec-gpu_b200 -- B200-native MSM engine behind the host API of kriptohaberciniz/0g-ec-gpu.

The directory name is not a Python identifier; import it with
``importlib.import_module("0g-ec-gpu_b200")`` or through the ``ec_gpu_b200`` shim at the repo root.

Contents:
  csrc/          hand-written sm_100a CUDA (field, curve, MSM kernels) + the C ABI (engine.cu)
  libmsm_b200.so built in-tree by ``make -C csrc`` / ``__graft_entry__.build()``
  _lib.py        ctypes binding of include/msm_b200.h (fails loudly when the library is missing)
  multiexp.py    mirror of ag_cuda_ec::multiexp (upload_multiexp_bases_*, multiple_multiexp_*)
  kernel.py      mirror of ec_gpu_proxy::multiexp::MultiexpKernel
  ec_fft.py      mirror of ag_cuda_ec::ec_fft (radix_ec_fft_*)
  fft.py         mirror of ec_gpu_proxy::fft::FftKernel (scalar-field FFT) and ec_gpu_proxy::ec_fft
"""
from ._lib import (  # noqa: F401
    BLS12_381_G1,
    BLS12_381_G2,
    BN254_G1,
    BN254_G2,
    CudaError,
    EcError,
    EcErrorAborted,
    EcErrorGpuTools,
    EcErrorSimple,
    build_library,
    describe_plan,
    fq_bytes,
    library_path,
    pipeline_shape,
    load_library,
)
from .multiexp import (  # noqa: F401
    DeviceData,
    Workspace,
    init_global_workspace,
    init_local_workspace,
    multiple_multiexp,
    multiple_multiexp_montgomery,
    multiple_multiexp_mt,
    multiple_multiexp_st,
    upload_multiexp_bases,
    upload_multiexp_bases_mt,
    upload_multiexp_bases_st,
)
from .kernel import MultiexpKernel, Worker  # noqa: F401
from .ec_fft import radix_ec_fft, radix_ec_fft_mt, radix_ec_fft_st  # noqa: F401
from .fft import EcFftKernel, FftKernel  # noqa: F401
from .sharding import chunk_size, shard_range  # noqa: F401
