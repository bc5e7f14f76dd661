"""Full-size bit-exact parity: the engine at the sizes BASELINE.json quotes, against tests/golden/fullsize.json
(the CPU oracle's results on the same seeded inputs, made by tests/golden/make_fullsize.py).

Parity notion: canonical affine coordinates after into_affine() on both sides
(ec-gpu-proxy/tests/multiexp.rs:38-105, line 99).  Every case runs the plain resident copy (table policy off)
AND the path a drop-in caller gets: upload_multiexp_bases -> multiple_multiexp twice, the second call
building the window table by policy (include/msm_b200.h, "Window tables by policy").
"""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from util import FQ, SEED

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "fullsize.json")) as f:
        g = json.load(f)
    assert g["seed"] == SEED
    return g


def _hex_points(oracle, curve, jac):
    fq = FQ[curve]
    xy, inf = oracle.to_affine(curve, np.ascontiguousarray(jac, dtype=np.uint8).reshape(-1, 3 * fq))
    return [{"x": bytes(r[:fq][::-1]).hex(), "y": bytes(r[fq:][::-1]).hex(), "inf": int(i)} for r, i in zip(xy, inf)]


def _digest(oracle, curve, jac):
    xy, inf = oracle.to_affine(curve, np.ascontiguousarray(jac, dtype=np.uint8).reshape(-1, 3 * FQ[curve]))
    return hashlib.sha256(xy.tobytes() + inf.tobytes()).hexdigest()


class _Synth:
    """Device-generated inputs of the synthetic stream (the generator is the oracle's: checked on a head
    and a tail window here, over 2^14 points in test_gpu_parity.py)."""

    def __init__(self, engine, oracle, curve, n_points, n_scalars):
        self.lib = engine.load_library()
        self.engine, self.curve, self.n_points, self.n_scalars = engine, curve, n_points, n_scalars
        self.ws = engine.Workspace(curve)
        h, fq = self.ws.handle, FQ[curve]
        self.dp, self.ds = ctypes.c_void_p(), ctypes.c_void_p()
        assert self.lib.msm_device_alloc(h, n_points * 2 * fq, ctypes.byref(self.dp)) == 0
        assert self.lib.msm_device_alloc(h, n_scalars * 32, ctypes.byref(self.ds)) == 0
        assert self.lib.msm_synth_points_device(h, SEED, 0, n_points, self.dp) == 0
        assert self.lib.msm_synth_scalars_device(h, SEED, 0, n_scalars, self.ds) == 0
        self.points = np.zeros((n_points, 2 * fq), dtype=np.uint8)
        self.scalars = np.zeros((n_scalars, 32), dtype=np.uint8)
        assert self.lib.msm_memcpy_d2h(h, self.points.ctypes.data, self.dp, self.points.nbytes) == 0
        assert self.lib.msm_memcpy_d2h(h, self.scalars.ctypes.data, self.ds, self.scalars.nbytes) == 0
        m = 512
        assert (self.points[:m] == oracle.gen_points(curve, SEED, m)).all()
        assert (self.points[-m:] == oracle.gen_points(curve, SEED, m, start=n_points - m)).all()
        assert (self.scalars[:m] == oracle.gen_scalars(curve, SEED, m)).all()
        assert (self.scalars[-m:] == oracle.gen_scalars(curve, SEED, m, start=n_scalars - m)).all()
        self.lib.msm_device_free(h, self.dp)
        self.dp = None

    def close(self):
        self.lib.msm_device_free(self.ws.handle, self.ds)
        self.ws.close()


def _run_paths(engine, s, num_chunks, expect_table_c=None):
    """plain resident copy, then the drop-in sequence; returns {path: results}."""
    lib, ws = s.lib, s.ws
    out = {}
    bases = engine.upload_multiexp_bases(ws, s.points)          # the reference API's upload
    sc = s.scalars
    assert lib.msm_bases_set_table_policy(ws.handle, bases._h, 0) == 0
    out["plain"] = engine.multiple_multiexp(ws, bases, sc, num_chunks, 8, True)
    assert lib.msm_bases_table_window(bases._h) == 0
    assert lib.msm_bases_set_table_policy(ws.handle, bases._h, 1) == 0
    out["first_call"] = engine.multiple_multiexp(ws, bases, sc, num_chunks, 8, True)   # still plain
    out["second_call"] = engine.multiple_multiexp(ws, bases, sc, num_chunks, 8, True)  # builds + uses the table
    t = ws.timings()
    assert lib.msm_bases_table_window(bases._h) != 0, "the second call of one shape builds the window table"
    if expect_table_c is not None:
        assert t["window_bits"] == expect_table_c, t
    # device-resident scalars, table in place
    n_out = out["second_call"].shape[0]
    do = ctypes.c_void_p()
    assert lib.msm_device_alloc(ws.handle, out["second_call"].nbytes, ctypes.byref(do)) == 0
    assert lib.msm_multiple_multiexp_device(ws.handle, bases._h, s.ds, s.n_scalars, num_chunks, do) == 0
    dev = np.zeros_like(out["second_call"])
    assert lib.msm_memcpy_d2h(ws.handle, dev.ctypes.data, do, dev.nbytes) == 0
    lib.msm_device_free(ws.handle, do)
    out["device_scalars"] = dev
    assert dev.shape[0] == n_out
    bases.free()
    return out


def test_bn254_2pow24_bit_exact(engine, oracle, golden):
    """BASELINE.json configs[2]: BN254 G1, 2^24 points -- the size the metric is quoted on."""
    s = _Synth(engine, oracle, 0, 1 << 24, 1 << 24)
    try:
        want = golden["bn254_2p24"]["result"]
        for path, got in _run_paths(engine, s, 1, expect_table_c=22).items():
            assert _hex_points(oracle, 0, got)[0] == want, path
    finally:
        s.close()


def test_bn254_2pow20_bit_exact(engine, oracle, golden):
    """BASELINE.json configs[1]: BN254 G1, 2^20 points on one GPU."""
    s = _Synth(engine, oracle, 0, 1 << 20, 1 << 20)
    try:
        want = golden["bn254_2p20"]["result"]
        for path, got in _run_paths(engine, s, 1).items():
            assert _hex_points(oracle, 0, got)[0] == want, path
    finally:
        s.close()


def test_bls381_2pow22_bit_exact(engine, oracle, golden):
    """BASELINE.json configs[3]: BLS12-381 G1, 2^22 points."""
    s = _Synth(engine, oracle, 1, 1 << 22, 1 << 22)
    try:
        want = golden["bls12_381_2p22"]["result"]
        for path, got in _run_paths(engine, s, 1).items():
            assert _hex_points(oracle, 1, got)[0] == want, path
    finally:
        s.close()


def test_batched_1024x4096_bit_exact(engine, oracle, golden):
    """BASELINE.json configs[4]: 1024 BN254 MSMs of 2^12 points (ag-cuda-ec/benches/multiexp.rs:19-22,56)."""
    g = golden["bn254_batched_1024x4096"]
    s = _Synth(engine, oracle, 0, g["L"], g["L"])
    try:
        for path, got in _run_paths(engine, s, g["num_chunks"]).items():
            pts = _hex_points(oracle, 0, got)
            bad = [i for i, (a, b) in enumerate(zip(pts, g["results"])) if a != b]
            assert not bad, (path, bad[:8])
            assert _digest(oracle, 0, got) == g["sha256"], path
    finally:
        s.close()


def test_amt_shape_bit_exact(engine, oracle, golden):
    """The AMT shape, 10 lines x 2^21 points, 2048 chunks (ag-cuda-ec/benches/amt.rs:18-55): 20480 results,
    compared through their SHA-256 and the first eight points."""
    g = golden["bn254_amt_10x2p21_2048"]
    s = _Synth(engine, oracle, 0, g["lines"] * g["L"], g["L"])
    try:
        for path, got in _run_paths(engine, s, g["num_chunks"]).items():
            assert got.shape[0] == g["lines"] * g["num_chunks"]
            assert _hex_points(oracle, 0, got[:8]) == g["first_results"], path
            assert _digest(oracle, 0, got) == g["sha256"], path
    finally:
        s.close()


def test_bls381_batched_256x4096_bit_exact(engine, oracle, golden):
    """The batched shape of configs[4] on the other curve: 256 BLS12-381 MSMs of 2^12 points."""
    g = golden["bls12_381_batched_256x4096"]
    s = _Synth(engine, oracle, 1, g["L"], g["L"])
    try:
        for path, got in _run_paths(engine, s, g["num_chunks"]).items():
            assert got.shape[0] == g["num_chunks"]
            assert _hex_points(oracle, 1, got[:8]) == g["first_results"], path
            assert _digest(oracle, 1, got) == g["sha256"], path
    finally:
        s.close()


@pytest.mark.parametrize("name,curve", [("bn254_g2_2p20", 2), ("bls12_381_g2_2p18", 3)])
def test_g2_fullsize_bit_exact(engine, oracle, golden, name, curve):
    """G2 over Fq2 at 2^20 / 2^18 points: same paths, same parity notion."""
    g = golden[name]
    assert g["curve"] == curve
    s = _Synth(engine, oracle, curve, g["n"], g["n"])
    try:
        for path, got in _run_paths(engine, s, 1).items():
            assert _hex_points(oracle, curve, got)[0] == g["result"], path
    finally:
        s.close()
