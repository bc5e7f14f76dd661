"""The CPU oracle against everything that can pin it in this container (no GPU needed):

  * public known answers (BN254 2G / 3G), Montgomery constants of SURVEY.md section 8c;
  * tests/golden/pyref_vectors.json   -- independent pure-Python big-integer arithmetic;
  * tests/golden/ref_cl_vectors.json  -- outputs of the reference's OWN device sources
    (ag-build/cl/*.cl) executed on the host (oracle/build_ref.py), committed as fixtures;
  * the same reference build run live when /root/reference is present.

The reference holds no golden vectors of its own and its Rust host code cannot be built here
(no Rust toolchain): see the header of oracle/msm_oracle.cpp.
"""
import json
import os

import numpy as np
import pytest

from util import FQ, assert_same_points

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {0: "bn254", 1: "bls12_381"}


def _load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def _b(hexstr, row):
    return np.frombuffer(bytes.fromhex(hexstr), dtype=np.uint8).reshape(-1, row).copy()


def _affine_ints(oracle, curve, jac):
    xy, inf = oracle.to_affine(curve, jac)
    out = []
    for i in range(len(xy)):
        if inf[i]:
            out.append(None)
        else:
            out.append([hex(int.from_bytes(xy[i, :FQ[curve]].tobytes(), "little")),
                        hex(int.from_bytes(xy[i, FQ[curve]:].tobytes(), "little"))])
    return out


@pytest.mark.parametrize("curve", [0, 1])
def test_constants(oracle, pyref, curve):
    cv = pyref.CURVES[curve]
    as_int = lambda a: int.from_bytes(a.tobytes(), "little")  # noqa: E731
    assert as_int(oracle.constant(curve, 0)) == cv.p
    assert as_int(oracle.constant(curve, 1)) == cv.R % cv.p
    assert as_int(oracle.constant(curve, 2)) == cv.R * cv.R % cv.p
    inv64 = as_int(oracle.constant(curve, 3))
    assert (inv64 * cv.p + 1) % (1 << 64) == 0
    assert as_int(oracle.constant(curve, 5)) == cv.r
    # SURVEY.md section 8c table (32-bit INV = low half of the 64-bit one)
    want32 = {0: 0xE4866389, 1: 0xFFFCFFFD}[curve]
    assert inv64 & 0xFFFFFFFF == want32
    gen = oracle.constant(curve, 4)
    assert cv.affine_from_bytes(gen.tobytes()) == cv.g
    assert cv.on_curve(cv.g) and oracle.on_curve(curve, gen)


def test_fp29_constants(pyref):
    """Constants baked into csrc/fp29.cuh (tools/gen_fp29_constants.py)."""
    import re
    p = pyref.BN254.p
    src = open(os.path.join(os.path.dirname(HERE), "0g-ec-gpu_b200", "csrc", "fp29.cuh")).read()

    def table(name):
        m = re.search(r"constexpr uint32_t %s\(int i\) \{\s*constexpr uint32_t t\[N\] = \{([^}]*)\}" % name, src)
        limbs = [int(x.strip().rstrip("u"), 16) for x in m.group(1).split(",")]
        assert all(l < (1 << 29) for l in limbs)
        return sum(l << (29 * i) for i, l in enumerate(limbs))

    Rp, R = 1 << 261, 1 << 256
    assert table("P") == p
    assert table("ONE") == Rp % p
    assert table("CONV_IN") == Rp * Rp * pow(R, -1, p) % p
    assert table("CONV_OUT") == R % p
    inv = int(re.search(r"INV = (0x[0-9a-f]+)u;\s*// -p\^-1 mod 2\^29", src).group(1), 16)
    assert (inv * p + 1) % (1 << 29) == 0


def test_public_known_answers(oracle, pyref):
    cv = pyref.BN254
    one = np.frombuffer((1).to_bytes(32, "little"), dtype=np.uint8)
    gen = oracle.constant(0, 4)
    for k, want in ((2, pyref.BN254_2G), (3, pyref.BN254_3G)):
        sc = np.frombuffer(k.to_bytes(32, "little"), dtype=np.uint8)
        jac = oracle.scalar_mul(0, gen, sc)
        assert cv.jacobian_from_bytes(jac.tobytes()) == want
        assert cv.mul(k, cv.g) == want
    assert cv.jacobian_from_bytes(oracle.scalar_mul(0, gen, one).tobytes()) == cv.g


@pytest.mark.parametrize("curve", [0, 1])
def test_against_pyref_fixture(oracle, curve):
    g = _load("pyref_vectors.json")["curves"][NAMES[curve]]
    fq = FQ[curve]
    a, b = _b("".join(g["fq_mul"]["a_mont"]), fq), _b("".join(g["fq_mul"]["b_mont"]), fq)
    assert (oracle.fq_op(curve, 2, a, b) == _b("".join(g["fq_mul"]["ab_mont"]), fq)).all()
    assert (oracle.fq_op(curve, 0, a, b) == _b("".join(g["fq_mul"]["a_plus_b_mont"]), fq)).all()
    assert (oracle.fq_op(curve, 1, a, b) == _b("".join(g["fq_mul"]["a_minus_b_mont"]), fq)).all()
    gen = oracle.constant(curve, 4)
    for k, want in g["kG"].items():
        sc = np.frombuffer(int(k).to_bytes(32, "little"), dtype=np.uint8)
        assert _affine_ints(oracle, curve, oracle.scalar_mul(curve, gen, sc).reshape(1, -1)) == [want]
    syn = g["synthetic"]
    n = len(syn["scalars"]) // 64
    assert oracle.gen_scalars(curve, syn["seed"], n).tobytes().hex() == syn["scalars"]
    assert oracle.gen_points(curve, syn["seed"], n).tobytes().hex() == syn["points_mont"]
    for case in g["msm"]:
        sc, pts = _b(case["scalars"], 32), _b(case["points_mont"], 2 * fq)
        # identity bases: the reference's CPU path errors (multiexp_cpu.rs:57-61); its GPU semantics
        # (and arkworks msm) treat them as contributing nothing -> multiple_multiexp with 1 chunk
        got = oracle.multiple_multiexp(curve, pts, sc, 1)
        assert _affine_ints(oracle, curve, got) == [case["result_affine"]], case["name"]
        if not (pts.reshape(len(pts), -1) == 0).all(axis=1).any():
            got2 = oracle.multiexp_cpu(curve, pts, sc).reshape(1, -1)
            assert _affine_ints(oracle, curve, got2) == [case["result_affine"]], case["name"]
            assert _affine_ints(oracle, curve, oracle.msm_naive(curve, pts, sc).reshape(1, -1)) == [case["result_affine"]]
    mm = g["multiple_multiexp"]
    got = oracle.multiple_multiexp(curve, _b(mm["points_mont"], 2 * fq), _b(mm["scalars"], 32), mm["chunks"])
    assert _affine_ints(oracle, curve, got) == mm["results_affine"]


def test_identity_base_is_an_error_in_multiexp_cpu(oracle):
    pts = oracle.gen_points(0, 5, 40)
    sc = oracle.gen_scalars(0, 5, 40)
    pts[7] = 0
    with pytest.raises(ValueError, match="identity element"):
        oracle.multiexp_cpu(0, pts, sc)
    sc[7] = 0  # skipped before the base is looked at (multiexp_cpu.rs:282-284)
    oracle.multiexp_cpu(0, pts, sc)


def test_window_choice(oracle):
    """c = 3 below 32 terms, ceil(ln n) otherwise (multiexp_cpu.rs:353-357)."""
    assert oracle.window_for(31) == 3
    assert oracle.window_for(32) == 4
    assert oracle.window_for(1 << 16) == 12
    assert oracle.window_for(1 << 24) == 17


@pytest.mark.parametrize("curve", [0, 1])
def test_against_reference_cl_fixture(oracle, curve):
    """Oracle == outputs of the reference's own ag-build/cl sources (committed fixture)."""
    g = _load("ref_cl_vectors.json")["curves"][NAMES[curve]]
    fq = FQ[curve]
    a, b = _b(g["fq"]["a"], fq), _b(g["fq"]["b"], fq)
    for op, nm in enumerate(["add", "sub", "mul", "sqr", "double", "mont", "unmont"]):
        assert (oracle.fq_op(curve, op, a, b) == _b(g["fq"]["ops"][nm], fq)).all(), nm
    pts = _b(g["ec"]["affine"], 2 * fq)
    one = oracle.constant(curve, 1)
    lifted = np.zeros((len(pts), 3 * fq), dtype=np.uint8)
    lifted[:, :2 * fq] = pts
    lifted[:, 2 * fq:] = one
    dbl = oracle.ec_op(curve, 2, lifted)
    assert_same_points(oracle, curve, dbl, _b(g["ec"]["double"], 3 * fq), "double")
    # same formulas (dbl-2009-l / madd-2007-bl / add-2007-bl) -> even the Jacobian bytes agree
    assert (dbl == _b(g["ec"]["double"], 3 * fq)).all()
    trip = oracle.ec_op(curve, 1, dbl, pts)
    assert (trip == _b(g["ec"]["double_plus_affine"], 3 * fq)).all()
    assert (oracle.ec_op(curve, 0, dbl, trip) == _b(g["ec"]["add_double_triple"], 3 * fq)).all()
    me = g["multiexp"]
    sc, bp = _b(me["scalars"], 32), _b(me["points_mont"], 2 * fq)
    want = oracle.multiple_multiexp(curve, bp, sc, me["chunks"])
    for run in me["runs"]:
        assert_same_points(oracle, curve, want, _b(run["results_jacobian"], 3 * fq),
                           f"POINT_multiexp w={run['window_size']} neg={run['neg_is_cheap']}")


@pytest.mark.parametrize("curve", [0, 1])
def test_against_reference_cl_live(oracle, curve):
    """Same comparison with the reference's sources compiled and run now (skipped on the GPU box,
    where /root/reference does not exist)."""
    from oracle import ref_cl

    if not os.path.isdir("/root/reference/ag-build/cl"):
        pytest.skip("reference sources not present")
    fq = FQ[curve]
    rng = np.random.default_rng(31 + curve)
    p = int.from_bytes(oracle.constant(curve, 0).tobytes(), "little")
    n = 400
    mk = lambda: np.frombuffer(b"".join((int.from_bytes(rng.bytes(fq + 8), "little") % p).to_bytes(fq, "little")  # noqa: E731
                                        for _ in range(n)), dtype=np.uint8).copy()
    a, b = mk(), mk()
    for op in range(7):
        assert (ref_cl.fq_op(curve, op, a, b) == oracle.fq_op(curve, op, a, b)).all(), op
    # the reference test shape: 2 lines x 32 chunks x 64 points (ag-cuda-ec/src/multiexp.rs:97-101)
    pts = oracle.gen_points(curve, 3, 64 * 32 * 2)
    sc = oracle.gen_scalars(curve, 3, 64 * 32)
    want = oracle.multiple_multiexp(curve, pts, sc, 32)
    for w, neg in ((4, True), (6, False)):
        assert_same_points(oracle, curve, ref_cl.multiple_multiexp(curve, pts, sc, 32, w, neg), want, f"w={w}")


# ---------------------------------------------------------------------------------------------
# EC-FFT: the oracle's restatement of serial_ec_fft
# (ec-gpu-proxy/src/ec_fft_cpu.rs:12-57) against the transform's definition evaluated by pyref.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("curve", [0, 1])
def test_ec_fft_oracle_vs_naive_dft_golden(oracle, curve):
    for case in _load("ec_fft_vectors.json")["curves"][NAMES[curve]]:
        jac = _b(case["input_jacobian"], 3 * FQ[curve])
        omegas = _b(case["omegas_mont"], 32)
        assert jac.shape[0] == 1 << case["log_n"]
        got = oracle.ec_fft(curve, jac, omegas[0])
        assert _affine_ints(oracle, curve, got) == case["output_affine"], f"log_n={case['log_n']}"


@pytest.mark.parametrize("curve", [0, 1])
def test_ec_fft_oracle_inverse_round_trip(oracle, pyref, curve):
    """FFT with omega, then with omega^-1, is n times the input (ag-cuda-ec/benches/ec_fft.rs:88-106)."""
    cv = pyref.CURVES[curve]
    log_n, g = 5, {0: 5, 1: 7}[curve]
    n = 1 << log_n
    omega = pow(g, (cv.r - 1) // n, cv.r)
    mont = lambda v: np.frombuffer((v * (1 << 256) % cv.r).to_bytes(32, "little"), dtype=np.uint8)  # noqa: E731
    pts = oracle.gen_points(curve, 99, n)
    jac = np.zeros((n, 3 * FQ[curve]), dtype=np.uint8)
    jac[:, :2 * FQ[curve]] = pts
    jac[:, 2 * FQ[curve]:] = oracle.constant(curve, 1)
    fwd = oracle.ec_fft(curve, jac, mont(omega))
    back = oracle.ec_fft(curve, fwd, mont(pow(omega, -1, cv.r)))
    want = np.stack([oracle.scalar_mul(curve, pts[i], np.frombuffer(n.to_bytes(32, "little"), dtype=np.uint8))
                     for i in range(n)])
    assert_same_points(oracle, curve, back, want, "ifft(fft(x)) == n x")


# ---------------------------------------------------------------------------------------------
# Scalar-field FFT: the oracle's restatement of serial_fft
# (ec-gpu-proxy/src/fft_cpu.rs:10-52) against the transform's definition in Python integers.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("curve", [0, 1])
def test_fr_fft_oracle_vs_naive_dft_golden(oracle, curve):
    for case in _load("fr_fft_vectors.json")["curves"][NAMES[curve]]:
        a = _b(case["input_mont"], 32)
        want = _b(case["output_mont"], 32)
        got = oracle.fr_fft(curve, a, _b(case["omega_mont"], 32)[0])
        assert (got == want).all(), f"log_n={case['log_n']}"


# ---------------------------------------------------------------------------------------------
# FFT / EC-FFT against the REFERENCE'S OWN kernels (ag-build/cl/fft.cl, ag-build/cl/ec-fft.cl) run on the
# host by oracle/build_ref.py under restatements of their host pass loops (ec-gpu-proxy/src/fft.rs:50-136,
# ag-cuda-ec/src/ec_fft.rs:13-99): committed fixture, and live when /root/reference is present.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("curve", [0, 1])
def test_fft_oracle_vs_reference_cl_fixture(oracle, curve):
    g = _load("ref_cl_fft_vectors.json")["curves"][NAMES[curve]]
    for case in g["fr_fft"]:
        got = oracle.fr_fft(curve, _b(case["input_mont"], 32), _b(case["omega_mont"], 32)[0])
        assert got.tobytes().hex() == case["output_mont"], f"fr_fft 2^{case['log_n']}"
    for case in g["ec_fft"]:
        jac = _b(case["input_jacobian"], 3 * FQ[curve])
        got = oracle.ec_fft(curve, jac, _b(case["omegas_mont"], 32)[0])
        xy, inf = oracle.to_affine(curve, got)
        assert xy.tobytes().hex() == case["output_affine_canonical"] and [int(v) for v in inf] == case["output_is_inf"]


@pytest.mark.parametrize("curve", [0, 1])
def test_fft_oracle_vs_reference_cl_live(oracle, pyref, curve):
    from oracle import ref_cl

    if not ref_cl.available():
        pytest.skip("reference sources not present (GPU box): the committed fixture covers this")
    cv = pyref.CURVES[curve]
    gen = {0: 5, 1: 7}[curve]
    for log_n in (1, 4, 8, 10):  # 10: two passes (8 + 2 rounds) of the reference kernel
        n = 1 << log_n
        om = np.frombuffer((pow(gen, (cv.r - 1) // n, cv.r) * (1 << 256) % cv.r).to_bytes(32, "little"), dtype=np.uint8).copy()
        a = oracle.gen_scalars(curve, 600 + log_n, n)
        assert (ref_cl.fr_fft(curve, a, om) == oracle.fr_fft(curve, a, om)).all(), f"fr_fft 2^{log_n}"
    for log_n in (1, 3, 6):
        n = 1 << log_n
        omega = pow(gen, (cv.r - 1) // n, cv.r)
        oms = np.zeros((32, 32), dtype=np.uint8)
        for i in range(32):
            oms[i] = np.frombuffer((pow(omega, 1 << i, cv.r) * (1 << 256) % cv.r).to_bytes(32, "little"), dtype=np.uint8)
        jac = np.zeros((n, 3 * FQ[curve]), dtype=np.uint8)
        jac[:, :2 * FQ[curve]] = oracle.gen_points(curve, 700 + log_n, n)
        jac[:, 2 * FQ[curve]:] = oracle.constant(curve, 1)
        assert_same_points(oracle, curve, ref_cl.ec_fft(curve, jac, oms), oracle.ec_fft(curve, jac, oms[0]), f"ec_fft 2^{log_n}")


def test_fullsize_golden_file_is_the_oracle(oracle):
    """tests/golden/fullsize.json (what the GPU tests and bench.py compare full-size results with) is the oracle's
    output: the 2^20 entry is recomputed here (a few seconds); the larger entries come from the same script
    (tests/golden/make_fullsize.py) and nest the same input stream."""
    import json

    g = json.load(open(os.path.join(HERE, "golden", "fullsize.json")))
    n = 1 << 20
    assert g["seed"] == 0x0BADC0DE and g["bn254_2p20"]["n"] == n
    pts, sc = oracle.gen_points(0, g["seed"], n), oracle.gen_scalars(0, g["seed"], n)
    xy, inf = oracle.to_affine(0, oracle.multiexp_cpu(0, pts, sc))
    got = {"x": bytes(xy[0, :32][::-1]).hex(), "y": bytes(xy[0, 32:][::-1]).hex(), "inf": int(inf[0])}
    assert got == g["bn254_2p20"]["result"]
    # a chunk of the batched golden: task 7 of 1024 x 4096 is the MSM of points / scalars [7 * 4096, 8 * 4096)
    lo, hi = 7 * 4096, 8 * 4096
    xy, inf = oracle.to_affine(0, oracle.multiexp_cpu(0, pts[lo:hi], sc[lo:hi]))
    want = g["bn254_batched_1024x4096"]["results"][7]
    assert {"x": bytes(xy[0, :32][::-1]).hex(), "y": bytes(xy[0, 32:][::-1]).hex(), "inf": int(inf[0])} == want


@pytest.mark.parametrize("name,curve", [("bn254_g2_2p20", 2), ("bls12_381_g2_2p18", 3)])
def test_fullsize_g2_golden_by_linearity(oracle, name, curve):
    """The G2 entries of fullsize.json, re-derived another way: MSM(first half) + MSM(second half) with a
    different thread split, compared after into_affine()."""
    import json

    g = json.load(open(os.path.join(HERE, "golden", "fullsize.json")))[name]
    n, fq = g["n"], oracle.FQ_BYTES[curve]
    pts, sc = oracle.gen_points(curve, 0x0BADC0DE, n), oracle.gen_scalars(curve, 0x0BADC0DE, n)
    a = oracle.multiexp_cpu(curve, pts[: n // 2], sc[: n // 2], nthreads=3)
    b = oracle.multiexp_cpu(curve, pts[n // 2:], sc[n // 2:], nthreads=5)
    xy, inf = oracle.to_affine(curve, oracle.ec_op(curve, 0, a, b))
    assert {"x": bytes(xy[0, :fq][::-1]).hex(), "y": bytes(xy[0, fq:][::-1]).hex(), "inf": int(inf[0])} == g["result"]


def test_fullsize_bls_batched_golden_chunk(oracle):
    """Task 5 of the 256 x 4096 BLS12-381 golden is the MSM of points / scalars [5 * 4096, 6 * 4096)."""
    import json

    g = json.load(open(os.path.join(HERE, "golden", "fullsize.json")))["bls12_381_batched_256x4096"]
    lo, hi = 5 * 4096, 6 * 4096
    pts, sc = oracle.gen_points(1, 0x0BADC0DE, 4096, start=lo), oracle.gen_scalars(1, 0x0BADC0DE, 4096, start=lo)
    xy, inf = oracle.to_affine(1, oracle.multiexp_cpu(1, pts, sc))
    assert {"x": bytes(xy[0, :48][::-1]).hex(), "y": bytes(xy[0, 48:][::-1]).hex(), "inf": int(inf[0])} == g["first_results"][5]
