"""Scalar-field FFT on the GPU through the C ABI / the FftKernel mirror,
bit-exact against the committed naive-DFT fixture and the oracle's restatement of serial_fft.  Mirrors
the reference's own tests (ec-gpu-proxy/src/fft.rs tests: gpu radix_fft vs serial_fft / parallel_fft)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {0: "bn254", 1: "bls12_381"}
GEN = {0: 5, 1: 7}
R = 1 << 256


@pytest.fixture(scope="module")
def kernels(engine):
    return {c: engine.FftKernel.create([0], c) for c in (0, 1)}


def _mont(v, r):
    return np.frombuffer((v * R % r).to_bytes(32, "little"), dtype=np.uint8).copy()


def _random_fr(curve, pyref, n, seed):
    r = pyref.CURVES[curve].r
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 40), dtype=np.uint8)
    vals = [int.from_bytes(raw[i].tobytes(), "little") % r for i in range(n)]
    return np.stack([np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8) for v in vals]).copy()  # canonical < r: valid Montgomery residues


@pytest.mark.parametrize("curve", [0, 1])
def test_fr_fft_golden(kernels, curve):
    with open(os.path.join(HERE, "golden", "fr_fft_vectors.json")) as f:
        cases = json.load(f)["curves"][NAMES[curve]]
    for case in cases:
        a = np.frombuffer(bytes.fromhex(case["input_mont"]), dtype=np.uint8).reshape(-1, 32).copy()
        om = np.frombuffer(bytes.fromhex(case["omega_mont"]), dtype=np.uint8).copy()
        kernels[curve].radix_fft(a, om, case["log_n"])
        assert a.tobytes().hex() == case["output_mont"], case["log_n"]


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9, 10, 11, 12, 14, 15, 16, 19, 20, 21])
def test_fr_fft_vs_oracle(kernels, oracle, pyref, curve, log_n):
    """Every pass structure: shared-memory-only (<= 10), one to three tile passes of 1..5 rounds."""
    n = 1 << log_n
    r = pyref.CURVES[curve].r
    a = _random_fr(curve, pyref, n, 1000 + log_n)
    om = _mont(pow(GEN[curve], (r - 1) // max(n, 2), r), r)
    want = oracle.fr_fft(curve, a, om)
    got = a.copy()
    kernels[curve].radix_fft(got, om, log_n)
    assert (got == want).all()


@pytest.mark.parametrize("curve", [0, 1])
def test_fr_fft_inverse_round_trip(kernels, pyref, curve):
    log_n = 13
    n = 1 << log_n
    r = pyref.CURVES[curve].r
    omega = pow(GEN[curve], (r - 1) // n, r)
    a = _random_fr(curve, pyref, n, 77)
    work = a.copy()
    kernels[curve].radix_fft(work, _mont(omega, r), log_n)
    kernels[curve].radix_fft(work, _mont(pow(omega, -1, r), r), log_n)
    for i in (0, 1, n // 2, n - 1):
        x = int.from_bytes(a[i].tobytes(), "little")
        assert int.from_bytes(work[i].tobytes(), "little") == x * n % r


def test_fr_fft_many_and_errors(engine, kernels, oracle, pyref):
    r = pyref.CURVES[0].r
    ins = [_random_fr(0, pyref, 1 << k, k) for k in (3, 6)]
    oms = [_mont(pow(5, (r - 1) // (1 << k), r), r) for k in (3, 6)]
    want = [oracle.fr_fft(0, a, om) for a, om in zip(ins, oms)]
    kernels[0].radix_fft_many(ins, oms, [3, 6])
    assert all((g == w).all() for g, w in zip(ins, want))
    with pytest.raises(ValueError):
        kernels[0].radix_fft(ins[0], oms[0], 4)
    aborting = engine.FftKernel.create_with_abort([0], lambda: True, 0)
    with pytest.raises(engine.EcErrorAborted):
        aborting.radix_fft(ins[0], oms[0], 3)


@pytest.mark.parametrize("curve", [0, 1])
def test_fr_fft_vs_reference_kernel_fixture(kernels, curve):
    """Bit-exact against outputs of the reference's own FIELD_radix_fft kernel (ag-build/cl/fft.cl) run on
    the host (tests/golden/ref_cl_fft_vectors.json, made by tests/golden/make_golden.py)."""
    with open(os.path.join(HERE, "golden", "ref_cl_fft_vectors.json")) as f:
        cases = json.load(f)["curves"][NAMES[curve]]["fr_fft"]
    for case in cases:
        a = np.frombuffer(bytes.fromhex(case["input_mont"]), dtype=np.uint8).reshape(-1, 32).copy()
        om = np.frombuffer(bytes.fromhex(case["omega_mont"]), dtype=np.uint8).copy()
        kernels[curve].radix_fft(a, om, case["log_n"])
        assert a.tobytes().hex() == case["output_mont"], case["log_n"]
