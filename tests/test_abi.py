"""The C-ABI library loads and exports every symbol include/msm_b200.h declares; without a GPU the
entry points fail loudly (no compute is attempted here, and there is no CPU fallback to reach)."""
import ctypes
import importlib
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(engine):
    L = importlib.import_module("0g-ec-gpu_b200._lib")
    lib = engine.load_library()
    names = L.header_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.msm_version()


def test_library_is_sm100a_only():
    """One architecture, no PTX-only fallback, no other vendors: cuobjdump lists sm_100a cubins."""
    import subprocess
    so = os.path.join(ROOT, "0g-ec-gpu_b200", "libmsm_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = {line.split(".")[-2] for line in out.splitlines() if ".cubin" in line}
    assert archs == {"sm_100a"}, out


def test_no_oracle_or_cpu_path_in_product():
    """The product package never imports, links or executes anything under oracle/."""
    pkg = os.path.join(ROOT, "0g-ec-gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, fn


def test_fails_loudly_without_gpu(engine):
    lib = engine.load_library()
    if lib.msm_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.CudaError):
        engine.Workspace(engine.BN254_G1)
    with pytest.raises(engine.EcErrorSimple, match="No working GPUs found"):
        engine.MultiexpKernel.create(None, engine.BN254_G1)
    h = ctypes.c_void_p()
    assert lib.msm_ctx_create(0, None, 1, ctypes.byref(h)) == 5  # MSM_ERR_NO_DEVICE
    assert b"No working GPUs" in lib.msm_last_error(None)
    assert lib.msm_ctx_create(7, None, 1, ctypes.byref(h)) == 1  # unknown curve


def test_missing_library_is_an_import_error(tmp_path, monkeypatch):
    L = importlib.import_module("0g-ec-gpu_b200._lib")
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "_LIB", str(tmp_path / "libmsm_b200.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        L.load_library()
