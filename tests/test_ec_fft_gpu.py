"""EC-FFT on the GPU through the C ABI / the radix_ec_fft mirror, against
the committed naive-DFT fixture and the oracle's restatement of serial_ec_fft.  Mirrors the reference's
own test (ag-cuda-ec/src/ec_fft.rs:113-145: degrees 4..8, compare with Radix2EvaluationDomain::fft)."""
import json
import os

import numpy as np
import pytest

from util import FQ, assert_same_points

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {0: "bn254", 1: "bls12_381"}
GEN = {0: 5, 1: 7}  # multiplicative generators of Fr that arkworks' FftField uses


@pytest.fixture(scope="module")
def ws(engine):
    return {c: engine.Workspace(c) for c in (0, 1)}


def _omegas(pyref, curve, n, inverse=False):
    cv = pyref.CURVES[curve]
    omega = pow(GEN[curve], (cv.r - 1) // n, cv.r)
    if inverse:
        omega = pow(omega, -1, cv.r)
    out = np.zeros((32, 32), dtype=np.uint8)
    for i in range(32):  # omegas[i] = omegas[i-1].square(), ag-cuda-ec/src/ec_fft.rs:121-124
        out[i] = np.frombuffer((pow(omega, 1 << i, cv.r) * (1 << 256) % cv.r).to_bytes(32, "little"), dtype=np.uint8)
    return out


def _lift(oracle, curve, pts):
    jac = np.zeros((pts.shape[0], 3 * FQ[curve]), dtype=np.uint8)
    jac[:, :2 * FQ[curve]] = pts
    jac[:, 2 * FQ[curve]:] = oracle.constant(curve, 1)
    return jac


@pytest.mark.parametrize("curve", [0, 1])
def test_ec_fft_golden(engine, oracle, ws, curve):
    with open(os.path.join(HERE, "golden", "ec_fft_vectors.json")) as f:
        cases = json.load(f)["curves"][NAMES[curve]]
    for case in cases:
        jac = np.frombuffer(bytes.fromhex(case["input_jacobian"]), dtype=np.uint8).reshape(-1, 3 * FQ[curve]).copy()
        omegas = np.frombuffer(bytes.fromhex(case["omegas_mont"]), dtype=np.uint8).reshape(-1, 32).copy()
        engine.radix_ec_fft(ws[curve], jac, omegas)
        xy, inf = oracle.to_affine(curve, jac)
        for i, want in enumerate(case["output_affine"]):
            if want is None:
                assert inf[i]
            else:
                assert not inf[i]
                got = [hex(int.from_bytes(xy[i, :FQ[curve]].tobytes(), "little")),
                       hex(int.from_bytes(xy[i, FQ[curve]:].tobytes(), "little"))]
                assert got == want, (case["log_n"], i)


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("log_n", [0, 1, 2, 4, 5, 6, 7, 10])
def test_ec_fft_vs_oracle(engine, oracle, pyref, ws, curve, log_n):
    n = 1 << log_n
    jac = _lift(oracle, curve, oracle.gen_points(curve, 0xEC0FF7 + log_n, n))
    if n >= 16:
        jac[3, 2 * FQ[curve]:] = 0  # infinity
        jac[7] = jac[6]             # equal points
    omegas = _omegas(pyref, curve, max(n, 2))
    want = oracle.ec_fft(curve, jac, omegas[0])
    got = jac.copy()
    engine.radix_ec_fft(ws[curve], got, omegas)
    assert_same_points(oracle, curve, got, want, f"ec_fft 2^{log_n}")


@pytest.mark.parametrize("curve", [0, 1])
def test_ec_fft_inverse_round_trip(engine, oracle, pyref, ws, curve):
    """radix_ec_fft with the powers of omega^-1 undoes it up to the factor n (ag-cuda-ec/benches/ec_fft.rs:88-106)."""
    log_n = 8
    n = 1 << log_n
    pts = oracle.gen_points(curve, 4242, n)
    jac = _lift(oracle, curve, pts)
    work = jac.copy()
    engine.radix_ec_fft(ws[curve], work, _omegas(pyref, curve, n))
    engine.radix_ec_fft(ws[curve], work, _omegas(pyref, curve, n, inverse=True))
    k = np.frombuffer(n.to_bytes(32, "little"), dtype=np.uint8)
    want = np.stack([oracle.scalar_mul(curve, pts[i], k) for i in range(n)])
    assert_same_points(oracle, curve, work, want, "ifft(fft(x)) == n x")


def test_ec_fft_argument_errors(engine, oracle, pyref, ws):
    jac = _lift(oracle, 0, oracle.gen_points(0, 1, 12))  # not a power of two: the reference asserts
    with pytest.raises(AssertionError):
        engine.radix_ec_fft(ws[0], jac, _omegas(pyref, 0, 16))
    jac = _lift(oracle, 0, oracle.gen_points(0, 1, 16))
    with pytest.raises(engine.CudaError):  # fewer omegas than rounds
        engine.radix_ec_fft(ws[0], jac, _omegas(pyref, 0, 16)[:3])


def test_ec_fft_and_fft_in_concurrent_threads(engine, oracle, pyref):
    """ag-cuda-ec/benches/ec_fft.rs:62-110 (bench_ec_fft_parallel): radix_ec_fft_mt from several host
    threads at once, each on its thread-local workspace; here together with scalar-field FFTs."""
    import threading

    curve, log_n = 0, 6
    n = 1 << log_n
    r = pyref.CURVES[curve].r
    omegas = _omegas(pyref, curve, n)
    inputs = [_lift(oracle, curve, oracle.gen_points(curve, 900 + t, n)) for t in range(4)]
    want = [oracle.ec_fft(curve, j, omegas[0]) for j in inputs]
    errs = []

    def work(t):
        try:
            engine.radix_ec_fft_mt(inputs[t], omegas, curve)
            k = engine.FftKernel.create([0], curve)
            a = oracle.gen_scalars(curve, 50 + t, 1 << 12)
            om = np.frombuffer((pow(GEN[curve], (r - 1) >> 12, r) * (1 << 256) % r).to_bytes(32, "little"), dtype=np.uint8).copy()
            ref = oracle.fr_fft(curve, a, om)
            k.radix_fft(a, om, 12)
            assert (a == ref).all()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for t in range(4):
        assert_same_points(oracle, curve, inputs[t], want[t], f"thread {t}")


@pytest.mark.parametrize("curve", [0, 1])
def test_ec_fft_vs_reference_kernel_fixture(engine, oracle, ws, curve):
    """Same group elements as the reference's own POINT_radix_fft kernel (ag-build/cl/ec-fft.cl) run on the
    host under the pass loop of radix_ec_fft (tests/golden/ref_cl_fft_vectors.json)."""
    with open(os.path.join(HERE, "golden", "ref_cl_fft_vectors.json")) as f:
        cases = json.load(f)["curves"][NAMES[curve]]["ec_fft"]
    for case in cases:
        jac = np.frombuffer(bytes.fromhex(case["input_jacobian"]), dtype=np.uint8).reshape(-1, 3 * FQ[curve]).copy()
        omegas = np.frombuffer(bytes.fromhex(case["omegas_mont"]), dtype=np.uint8).reshape(-1, 32).copy()
        engine.radix_ec_fft(ws[curve], jac, omegas)
        xy, inf = oracle.to_affine(curve, jac)
        assert xy.tobytes().hex() == case["output_affine_canonical"], case["log_n"]
        assert [int(v) for v in inf] == case["output_is_inf"]
