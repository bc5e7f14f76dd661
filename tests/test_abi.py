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


def _header_prototypes():
    """name -> number of parameters, parsed from include/msm_b200.h"""
    import re

    text = open(os.path.join(ROOT, "include", "msm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(msm_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_ctypes_table_matches_header():
    """Every prototype of the header has a ctypes signature with the same number of parameters (a drifted
    binding would pass garbage through the ABI)."""
    import inspect
    import re

    L = importlib.import_module("0g-ec-gpu_b200._lib")
    protos = _header_prototypes()
    src = inspect.getsource(L.load_library)
    table = {}
    for m in re.finditer(r'"(msm_[a-z0-9_]+)":\s*\(\[(.*?)\],\s*[^\n]+\),?\n', src, flags=re.S):
        inner = m.group(2).strip()
        # arguments are comma-separated at bracket depth 0
        depth, n = 0, 1 if inner else 0
        for ch in inner:
            depth += ch in "([" 
            depth -= ch in ")]"
            n += ch == "," and depth == 0
        table[m.group(1)] = n
    missing = sorted(set(protos) - set(table))
    assert not missing, f"no ctypes signature for {missing}"
    wrong = {k: (protos[k], table[k]) for k in protos if protos[k] != table[k]}
    assert not wrong, f"parameter counts differ (header, ctypes): {wrong}"


def test_rust_ffi_declarations_match_header():
    """rust/msm-b200-sys declares a subset of the header with the same parameter counts."""
    import re

    protos = _header_prototypes()
    text = open(os.path.join(ROOT, "rust", "msm-b200-sys", "src", "lib.rs")).read()
    decls = re.findall(r"pub fn (msm_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]+)?;", text, flags=re.S)
    assert len(decls) >= 15
    for name, args in decls:
        assert name in protos, name
        n = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        assert n == protos[name], (name, n, protos[name])


def test_header_is_valid_c99_and_cpp(tmp_path):
    """The boundary is a C ABI: the header must compile as plain C (cgo / bindgen / a C caller) and as C++,
    and a C translation unit that calls through it must link against the library."""
    import subprocess

    hdr = os.path.join(ROOT, "include", "msm_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    src = tmp_path / "caller.c"
    src.write_text('#include "msm_b200.h"\n#include <stdio.h>\n'
                   "int main(void) { msm_ctx* c = 0; int rc = msm_ctx_create(MSM_CURVE_BN254_G1, 0, 1, &c);\n"
                   '  printf("%s rc=%d devices=%d\\n", msm_version(), rc, msm_device_count());\n'
                   "  if (rc == MSM_OK) msm_ctx_destroy(c);\n"
                   "  msm_plan_info p; unsigned n = 0; double g = 0;\n"  # the two host-only calls work from plain C, GPU or not
                   "  if (msm_plan_describe(MSM_CURVE_BN254_G1, (size_t)1 << 24, 1, 1, 22, 1, 2.0, &p) != MSM_OK) return 2;\n"
                   "  if (msm_pipeline_shape((size_t)1 << 24, 1, 1, 55.0f, 32.7f, &n, &g) != MSM_OK) return 3;\n"
                   '  printf("plan c=%u W=%u S=%u sub=%u growth=%.2f\\n", p.window_bits, p.num_windows, p.slice_len, n, g);\n'
                   "  return 0; }\n")
    exe = tmp_path / "caller"
    so_dir = os.path.join(ROOT, "0g-ec-gpu_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", so_dir, "-lmsm_b200", "-Wl,-rpath," + so_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "msm_b200" in out.stdout, out.stdout + out.stderr
    assert "plan c=22 W=12 S=333 sub=3 growth=2.85" in out.stdout, out.stdout
