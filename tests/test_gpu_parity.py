"""GPU parity: the CUDA engine, called through the C ABI, against the CPU oracle.

Integer arithmetic only: the bar is bit-exact equality of the canonical affine coordinates
(the reference compares after into_affine(), ec-gpu-proxy/tests/multiexp.rs:99).
"""
import ctypes
import json
import os

import numpy as np
import pytest

from util import FQ, SEED, adversarial_inputs, assert_same_points

pytestmark = pytest.mark.gpu

CURVES = [0, 1]


@pytest.fixture(scope="module")
def ws(engine):
    return {c: engine.Workspace(c) for c in CURVES}


def _rand_fq(oracle, curve, n, rng):
    p = int.from_bytes(oracle.constant(curve, 0).tobytes(), "little")
    vals = [int.from_bytes(rng.bytes(FQ[curve] + 8), "little") % p for _ in range(n)]
    vals[0], vals[1], vals[2] = 0, p - 1, 1
    return np.frombuffer(b"".join(v.to_bytes(FQ[curve], "little") for v in vals), dtype=np.uint8).copy()


@pytest.mark.parametrize("curve", CURVES)
def test_field_ops(engine, oracle, ws, curve):
    """Counterpart of ag-build/src/tests/test_fields.rs:11-107 (add, sub, mul, sqr, double, mont,
    unmont) plus inverse / neg, 4096 samples per op instead of 10."""
    lib = engine.load_library()
    rng = np.random.default_rng(1234 + curve)
    n = 4096
    a, b = _rand_fq(oracle, curve, n, rng), _rand_fq(oracle, curve, n, rng)
    for op in range(9):
        aa = a.copy()
        if op == 7:
            aa[: FQ[curve]] = a[FQ[curve]: 2 * FQ[curve]]  # no inverse of zero
        want = oracle.fq_op(curve, op, aa, b)
        got = np.zeros_like(aa)
        rc = lib.msm_test_fq_op(ws[curve].handle, op, aa.ctypes.data, b.ctypes.data, got.ctypes.data, n)
        assert rc == 0
        assert (got == want).all(), f"fq op {op} curve {curve}"


@pytest.mark.parametrize("curve", CURVES)
def test_ec_ops(engine, oracle, ws, curve):
    """Counterpart of ag-build/src/tests/test_ec.rs:7-37 for add / mixed add / double incl. the
    exceptional cases (equal inputs, inverse inputs, infinity)."""
    lib = engine.load_library()
    n, fq = 512, FQ[curve]
    pts = oracle.gen_points(curve, 7, n)
    sc = oracle.gen_scalars(curve, 9, n)
    jac = np.stack([oracle.scalar_mul(curve, pts[i], sc[i]) for i in range(n)])
    jac2 = np.stack([oracle.scalar_mul(curve, pts[(i * 7 + 3) % n], sc[(i + 1) % n]) for i in range(n)])
    jac2[0] = jac[0]
    jac2[2] = 0
    jac[3] = 0
    one = oracle.constant(curve, 1)
    lifted = np.zeros((n, 3 * fq), dtype=np.uint8)
    lifted[:, : 2 * fq] = pts
    lifted[:, 2 * fq:] = one
    neg = pts.copy()
    neg[:, fq:] = oracle.fq_op(curve, 8, pts[:, fq:].copy())
    for op, a, b in ((0, jac, jac2), (1, jac, pts), (2, jac, None), (1, lifted, pts), (1, lifted, neg)):
        want = oracle.ec_op(curve, op, a, b)
        got = np.zeros_like(a)
        rc = lib.msm_test_ec_op(ws[curve].handle, op, a.ctypes.data, None if b is None else b.ctypes.data,
                                got.ctypes.data, n)
        assert rc == 0
        assert_same_points(oracle, curve, got, want, f"ec op {op}")


def _synth(engine, ws, curve, n, seed=SEED, start=0):
    """Device-generated inputs copied back to the host."""
    lib = engine.load_library()
    h = ws.handle
    dp, ds = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.msm_device_alloc(h, n * 2 * FQ[curve], ctypes.byref(dp)) == 0
    assert lib.msm_device_alloc(h, n * 32, ctypes.byref(ds)) == 0
    assert lib.msm_synth_points_device(h, seed, start, n, dp) == 0
    assert lib.msm_synth_scalars_device(h, seed, start, n, ds) == 0
    pts = np.zeros((n, 2 * FQ[curve]), dtype=np.uint8)
    sc = np.zeros((n, 32), dtype=np.uint8)
    assert lib.msm_memcpy_d2h(h, pts.ctypes.data, dp, pts.nbytes) == 0
    assert lib.msm_memcpy_d2h(h, sc.ctypes.data, ds, sc.nbytes) == 0
    lib.msm_device_free(h, dp)
    lib.msm_device_free(h, ds)
    return pts, sc


@pytest.mark.parametrize("curve", CURVES)
def test_synthetic_inputs_match_oracle_generator(engine, oracle, ws, curve):
    n = 5000
    pts, sc = _synth(engine, ws[curve], curve, n, start=123)
    assert (sc == oracle.gen_scalars(curve, SEED, n, start=123)).all()
    assert (pts == oracle.gen_points(curve, SEED, n, start=123)).all()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("log_n", [0, 5, 10, 11, 14, 16])
def test_multiexp_vs_cpu(engine, oracle, curve, log_n):
    """ec-gpu-proxy/tests/multiexp.rs:38-105 (gpu_multiexp_consistency): MultiexpKernel::multiexp
    == multiexp_cpu after into_affine; sizes 2^10, 2^11 as there, plus smaller and larger."""
    n = 1 << log_n
    pts = oracle.gen_points(curve, SEED, n)
    sc = oracle.gen_scalars(curve, SEED, n)
    kern = engine.MultiexpKernel.create([0], curve)
    got = kern.multiexp(engine.Worker(), pts, sc, 0)
    want = oracle.multiexp_cpu(curve, pts, sc)
    assert_same_points(oracle, curve, got, want, f"n=2^{log_n}")


@pytest.mark.parametrize("curve", CURVES)
def test_multiexp_skip_and_empty(engine, oracle, curve):
    n = 300
    pts = oracle.gen_points(curve, SEED, n + 17)
    sc = oracle.gen_scalars(curve, SEED, n)
    kern = engine.MultiexpKernel.create([0], curve)
    got = kern.multiexp(engine.Worker(), pts, sc, 17)
    want = oracle.multiexp_cpu(curve, pts[17:], sc)
    assert_same_points(oracle, curve, got, want, "skip")
    empty = kern.multiexp(engine.Worker(), pts, sc[:0], 0)
    assert not empty[2 * FQ[curve]:].any()  # z == 0: infinity


@pytest.mark.parametrize("curve", CURVES)
def test_multiple_multiexp_batch(engine, oracle, ws, curve):
    """ag-cuda-ec/src/multiexp.rs:93-145 (test_multiexp_batch): 2 lines x 32 chunks x 64 points,
    window_size 1..=9 x neg_is_cheap in {true,false}; every combination must give the same result
    as the per-chunk CPU MSM."""
    CHUNK, CHUNKS, LINES = 64, 32, 2
    L = CHUNK * CHUNKS
    pts = oracle.gen_points(curve, SEED + 1, L * LINES)
    sc = oracle.gen_scalars(curve, SEED + 1, L)
    bases_gpu = engine.upload_multiexp_bases(ws[curve], pts)
    assert bases_gpu.size() == pts.nbytes
    want = oracle.multiple_multiexp(curve, pts, sc, CHUNKS)
    for window_size in range(1, 10):
        for neg in (True, False):
            got = engine.multiple_multiexp(ws[curve], bases_gpu, sc, CHUNKS, window_size, neg)
            assert got.shape == want.shape
            assert_same_points(oracle, curve, got, want, f"w={window_size} neg={neg}")


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("c", [2, 3, 5, 8, 11, 13, 16])
def test_result_independent_of_engine_window(engine, oracle, ws, curve, c):
    n = 3000
    pts, sc = adversarial_inputs(oracle, curve, n)
    w = ws[curve]
    w.set_window_bits(c)
    try:
        bases_gpu = engine.upload_multiexp_bases(w, pts)
        got = engine.multiple_multiexp(w, bases_gpu, sc, 1, 8, True)
        assert w.timings()["window_bits"] == c
    finally:
        w.set_window_bits(0)
    want = oracle.multiple_multiexp(curve, pts, sc, 1)
    assert_same_points(oracle, curve, got, want, f"c={c}")


@pytest.mark.parametrize("curve", CURVES)
def test_edge_cases(engine, oracle, ws, curve):
    """Zero / one / r-1 scalars, identity bases, repeated bases, P and -P."""
    n = 4096
    pts, sc = adversarial_inputs(oracle, curve, n)
    bases_gpu = engine.upload_multiexp_bases(ws[curve], pts)
    for chunks in (1, 4, 64):
        got = engine.multiple_multiexp(ws[curve], bases_gpu, sc, chunks, 8, True)
        want = oracle.multiple_multiexp(curve, pts, sc, chunks)
        assert_same_points(oracle, curve, got, want, f"adversarial chunks={chunks}")
    # all-zero scalars -> infinity; all-identity bases -> infinity
    z = np.zeros_like(sc)
    got = engine.multiple_multiexp(ws[curve], bases_gpu, z, 2, 8, True)
    assert not got[:, 2 * FQ[curve]:].any()
    ident = engine.upload_multiexp_bases(ws[curve], np.zeros_like(pts))
    got = engine.multiple_multiexp(ws[curve], ident, sc, 2, 8, True)
    assert not got[:, 2 * FQ[curve]:].any()
    # ragged: L not divisible by num_chunks drops the tail (ag-build/cl/multiexp.cl:235)
    got = engine.multiple_multiexp(ws[curve], bases_gpu, sc, 3, 8, True)
    want = oracle.multiple_multiexp(curve, pts, sc, 3)
    assert_same_points(oracle, curve, got, want, "ragged")


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("c", [0, 11, 14])
def test_window_table_path(engine, oracle, ws, curve, c):
    """msm_bases_precompute: the folded window table gives the same group element as the plain
    resident copy and as the oracle, including the edge cases; other call shapes on the same bases
    keep working."""
    n = 6000
    pts, sc = adversarial_inputs(oracle, curve, n)
    w = ws[curve]
    bases_gpu = engine.upload_multiexp_bases(w, pts)
    want = oracle.multiple_multiexp(curve, pts, sc, 1)
    plain = engine.multiple_multiexp(w, bases_gpu, sc, 1, 8, True)
    assert_same_points(oracle, curve, plain, want, "plain")
    tc = bases_gpu.precompute(c)
    assert tc >= 11 and (c == 0 or tc == c)
    folded = engine.multiple_multiexp(w, bases_gpu, sc, 1, 8, True)
    t = w.timings()
    assert t["window_bits"] == tc
    assert_same_points(oracle, curve, folded, want, "folded")
    # a chunked call on the same handle takes the ordinary path
    got = engine.multiple_multiexp(w, bases_gpu, sc, 8, 8, True)
    assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, pts, sc, 8), "chunked after precompute")


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("lines,chunks,chunk_len", [(1, 16, 300), (3, 8, 512), (2, 64, 64)])
def test_window_table_chunked_and_multiline(engine, oracle, ws, curve, lines, chunks, chunk_len):
    """SURVEY.md section 8f row 1 (AMT / per-segment shapes, ag-cuda-ec/benches/{multiexp,amt}.rs): a table
    built with msm_bases_precompute_chunked serves chunked and multi-line calls -- every task folds
    its windows into one bucket set -- and gives the same group elements as the plain path and as
    the oracle, on the adversarial input set (zero / one / r-1 scalars, identity bases, P and -P)."""
    L = chunks * chunk_len
    pts, sc = adversarial_inputs(oracle, curve, L * lines)
    sc = sc[:L]
    w = ws[curve]
    bases_gpu = engine.upload_multiexp_bases(w, pts)
    want = oracle.multiple_multiexp(curve, pts, sc, chunks)
    plain = engine.multiple_multiexp(w, bases_gpu, sc, chunks, 8, True)
    plain_c = w.timings()["window_bits"]
    assert_same_points(oracle, curve, plain, want, "plain")
    tc = bases_gpu.precompute_chunked(chunk_len)
    assert 8 <= tc <= 24
    folded = engine.multiple_multiexp(w, bases_gpu, sc, chunks, 8, True)
    t = w.timings()
    assert t["window_bits"] == tc and tc > plain_c, (tc, plain_c)
    assert folded.shape[0] == lines * chunks
    assert_same_points(oracle, curve, folded, want, "folded chunked")
    # a different chunking of the same row on the same table (cost model decides; result unchanged)
    got = engine.multiple_multiexp(w, bases_gpu, sc, chunks // 2, 8, True)
    assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, pts, sc, chunks // 2), "other chunking")
    # a shorter scalar row: more lines are inferred (points / L), table stride stays the shard size
    L2 = L // 2
    got = engine.multiple_multiexp(w, bases_gpu, sc[:L2], chunks // 2, 8, True)
    assert got.shape[0] == (L * lines // L2) * (chunks // 2)
    assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, pts, sc[:L2], chunks // 2), "short row")


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n_sub,table_c", [(2, 0), (3, 12), (8, 0), (8, 14), (5, 9)])
def test_pipelined_sub_batches(engine, oracle, ws, curve, n_sub, table_c, monkeypatch):
    """Host scalars of a single-task call are uploaded and processed in sub-batches that continue
    one shared bucket array (carry_in): same group element for any split, with and without a
    window table, on the adversarial set (so that buckets receive entries from several sub-batches,
    cancel to infinity in between, and stay empty in some sub-batches)."""
    monkeypatch.setenv("MSM_B200_PIPELINE", str(n_sub))
    n = 5003  # not a multiple of the split
    pts, sc = adversarial_inputs(oracle, curve, n)
    w = ws[curve]
    bases_gpu = engine.upload_multiexp_bases(w, pts)
    if table_c >= 11:
        bases_gpu.precompute(table_c)
    elif table_c:
        w.set_window_bits(table_c)
    try:
        got = engine.multiple_multiexp(w, bases_gpu, sc, 1, 8, True)
        t = w.timings()
    finally:
        w.set_window_bits(0)
    assert t["sub_batches"] == n_sub
    assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, pts, sc, 1), f"pipeline {n_sub}")


@pytest.mark.parametrize("groups", [0, 3, 8])
def test_task_groups_pipelined_upload(engine, oracle, ws, groups, monkeypatch):
    """A row of many independent tasks (the reference's bench geometry, ag-cuda-ec/benches/multiexp.rs:19-22) is
    uploaded in groups of whole tasks, each sorted and accumulated into its own bucket range as it lands, with one
    reduction over all tasks at the end (enqueue_msm, Plan::by_task): 37 tasks of
    28 411 points (+ a dropped tail, ag-build/cl/multiexp.cl:235) in 3 (default, and forced) and 8 groups, on the plain resident
    copy and on the window table the second call builds -- every task equal to the oracle's."""
    curve, chunks, chunk_len = 0, 37, 28411
    L = chunks * chunk_len + 5
    assert L >= 1 << 20
    if groups:
        monkeypatch.setenv("MSM_B200_PIPELINE", str(groups))
    w = ws[curve]
    pts, sc = _synth(engine, w, curve, L)
    want = oracle.multiple_multiexp(curve, pts, sc, chunks)
    bases = engine.upload_multiexp_bases(w, pts)
    lib = engine.load_library()
    for call in range(3):  # plain, plain + table build, table
        got = engine.multiple_multiexp(w, bases, sc, chunks, 8, True)
        assert w.timings()["sub_batches"] == (groups or 3)
        assert got.shape[0] == chunks
        assert_same_points(oracle, curve, got, want, f"task groups {groups}, call {call}")
    # the cost model may or may not want a table for this shape; with one forced, the groups index it by point_offset
    bases.precompute_chunked(chunk_len)
    assert lib.msm_bases_table_window(bases._h) != 0
    got = engine.multiple_multiexp(w, bases, sc, chunks, 8, True)
    assert_same_points(oracle, curve, got, want, f"task groups {groups}, explicit table")
    bases.free()


def test_sub_batch_growth_adapts_to_measured_speeds(engine, oracle, monkeypatch):
    """The first pipelined call of a shape splits 2^23 host scalars 4-fold with sizes doubling; once the workspace has
    measured its upload rate and the shape's device time, the split follows them (3 sub-batches growing up to 3-fold
    on a fast link, 4 growing more slowly on a slow one); MSM_B200_PIPELINE_STATIC=1 pins the first form.  Every form
    gives the same point."""
    curve, n = 0, 1 << 23
    lib = engine.load_library()
    w = engine.Workspace(curve)
    try:
        h, fq = w.handle, FQ[curve]
        dp, ds = ctypes.c_void_p(), ctypes.c_void_p()
        assert lib.msm_device_alloc(h, n * 2 * fq, ctypes.byref(dp)) == 0
        assert lib.msm_device_alloc(h, n * 32, ctypes.byref(ds)) == 0
        assert lib.msm_synth_points_device(h, SEED, 0, n, dp) == 0
        assert lib.msm_synth_scalars_device(h, SEED, 0, n, ds) == 0
        bh = ctypes.c_void_p()
        assert lib.msm_bases_from_device(h, dp, n, ctypes.byref(bh)) == 0
        lib.msm_device_free(h, dp)
        sc = np.zeros((n, 32), dtype=np.uint8)
        assert lib.msm_memcpy_d2h(h, sc.ctypes.data, ds, sc.nbytes) == 0
        lib.msm_device_free(h, ds)
        assert lib.msm_host_register(sc.ctypes.data, sc.nbytes) == 0
        try:
            outs, subs = [], []
            for call in range(4):  # plain, plain -> table built, table, table
                out = np.zeros((1, 3 * fq), dtype=np.uint8)
                assert lib.msm_multiple_multiexp(h, bh, sc.ctypes.data, n, 1, 8, 1, out.ctypes.data) == 0
                outs.append(out)
                subs.append(w.timings()["sub_batches"])
            assert subs[0] == 4, subs                      # nothing measured yet
            assert subs[1] == 4, subs                      # builds the table: no device time on the table yet
            assert all(s in (3, 4) for s in subs), subs
            monkeypatch.setenv("MSM_B200_PIPELINE_STATIC", "1")
            out = np.zeros((1, 3 * fq), dtype=np.uint8)
            assert lib.msm_multiple_multiexp(h, bh, sc.ctypes.data, n, 1, 8, 1, out.ctypes.data) == 0
            assert w.timings()["sub_batches"] == 4
            outs.append(out)
            for i, o in enumerate(outs[1:]):
                assert_same_points(oracle, curve, o, outs[0], f"call {i + 1} of {subs}")
            # and it is the oracle's point for this input (tests/golden/fullsize.json)
            want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.json")))["bn254_2p23"]["result"]
            xy, inf = oracle.to_affine(curve, outs[-1])
            assert bytes(xy[0, :fq][::-1]).hex() == want["x"] and bytes(xy[0, fq:][::-1]).hex() == want["y"] and not inf[0]
            print("sub-batches per call:", subs)
        finally:
            lib.msm_host_unregister(sc.ctypes.data)
            lib.msm_bases_free(bh)
    finally:
        w.close()


def test_window_table_sharded_resident(engine, oracle):
    """MultiexpKernel with resident sharded bases + tables (one device here; N devices in bench)."""
    lib = engine.load_library()
    n = 1 << 15
    pts, sc = oracle.gen_points(0, SEED, n), oracle.gen_scalars(0, SEED, n)
    kern = engine.MultiexpKernel.create([0], 0)
    res = kern.upload_bases(pts)
    assert lib.msm_bases_precompute(kern.workspace.handle, res, 0) == 0
    got = kern.multiexp_resident(res, sc, 0)
    assert_same_points(oracle, 0, got, oracle.multiexp_cpu(0, pts, sc), "resident + table")
    part = kern.multiexp_resident(res, sc[:1000], 5)  # a sub-range falls back to the plain copy
    assert_same_points(oracle, 0, part, oracle.multiexp_cpu(0, pts[5:], sc[:1000]), "sub-range")
    res.free()


@pytest.mark.parametrize("curve", CURVES)
def test_montgomery_scalars_on_device(engine, oracle, pyref, ws, curve):
    """exponents handed over in Montgomery form are converted on the
    device (FIELD_unmont, ag-build/cl/field.cl:365-377) and give the same MSM."""
    cv = pyref.CURVES[curve]
    n = 2048
    pts, sc = adversarial_inputs(oracle, curve, n)
    R = 1 << 256
    mont = np.zeros_like(sc)
    for i in range(n):
        k = int.from_bytes(sc[i].tobytes(), "little")
        mont[i] = np.frombuffer((k * R % cv.r).to_bytes(32, "little"), dtype=np.uint8)
    bases_gpu = engine.upload_multiexp_bases(ws[curve], pts)
    got = engine.multiple_multiexp_montgomery(ws[curve], bases_gpu, mont, 4)
    want = oracle.multiple_multiexp(curve, pts, sc, 4)
    assert_same_points(oracle, curve, got, want, "montgomery scalars")


@pytest.mark.parametrize("curve", CURVES)
def test_montgomery_scalars_fused_into_every_sort(engine, oracle, ws, curve, monkeypatch):
    """The Montgomery -> canonical conversion is fused into the digit decomposition (sort.cu: load_scalar_geo):
    every sort that reads the row must convert -- single-level, binned (forced), and the pipelined sub-batch
    upload of a 2^20 row, with and without the window table."""
    n = 1 << 20
    pts, sc = _synth(engine, ws[curve], curve, n)
    mont = oracle.fr_op(curve, 0, sc)  # canonical -> Montgomery
    want = oracle.multiexp_cpu(curve, pts, sc)
    bases = engine.upload_multiexp_bases(ws[curve], pts)
    for call in range(3):  # plain, plain, table
        got = engine.multiple_multiexp_montgomery(ws[curve], bases, mont, 1)
        assert ws[curve].timings()["sub_batches"] == 2
        assert_same_points(oracle, curve, got, want, f"montgomery row, call {call}")
    monkeypatch.setenv("MSM_B200_SORT", "binned")
    small = 1 << 14
    got = engine.multiple_multiexp_montgomery(ws[curve], bases, mont[:small], 8)
    assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, pts, sc[:small], 8), "binned sort, 64 lines x 8 chunks")
    bases.free()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("c,chunks", [(13, 1), (16, 1), (9, 16), (20, 1)])
def test_two_level_scatter_path(engine, oracle, ws, curve, c, chunks, monkeypatch):
    """The partition + final-scatter sort that large calls use, forced on a small adversarial
    input (skewed buckets, identity bases) for generic and compile-time window sizes."""
    monkeypatch.setenv("MSM_B200_PARTITION", "1")
    n = 5000
    pts, sc = adversarial_inputs(oracle, curve, n)
    w = ws[curve]
    w.set_window_bits(c)
    try:
        bases_gpu = engine.upload_multiexp_bases(w, pts)
        got = engine.multiple_multiexp(w, bases_gpu, sc, chunks, 8, True)
        assert w.timings()["scatter_passes"] == 0
    finally:
        w.set_window_bits(0)
    assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, pts, sc, chunks), f"partition c={c}")
    # and through a window table
    bases_gpu.precompute(14)
    got = engine.multiple_multiexp(w, bases_gpu, sc, 1, 8, True)
    assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, pts, sc, 1), "partition + table")


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("c,chunks,lines,n_sub", [(13, 1, 1, 1), (16, 1, 1, 3), (9, 16, 2, 1), (20, 1, 1, 1), (11, 4, 1, 1)])
def test_binned_sort_path(engine, oracle, ws, curve, c, chunks, lines, n_sub, monkeypatch):
    """The binned sort large calls use (coarse shared-memory histogram, partition, per-tile
    shared-memory histogram and placement), forced on small adversarial inputs: skewed buckets (many
    tiles in one bin), identity bases, several tasks and lines, pipelined sub-batches, and a window table."""
    monkeypatch.setenv("MSM_B200_SORT", "binned")
    monkeypatch.setenv("MSM_B200_PIPELINE", str(n_sub))
    n = 40000 // lines if chunks == 1 else 4096
    n -= n % chunks
    pts, sc = adversarial_inputs(oracle, curve, n * lines)
    sc = sc[:n]
    sc[n // 2: n // 2 + n // 4] = sc[20]  # a quarter of the scalars equal: a few very heavy buckets
    w = ws[curve]
    w.set_window_bits(c)
    try:
        bases_gpu = engine.upload_multiexp_bases(w, pts)
        got = engine.multiple_multiexp(w, bases_gpu, sc, chunks, 8, True)
        t = w.timings()
        assert t["scatter_passes"] == 0 and t["window_bits"] == c
    finally:
        w.set_window_bits(0)
    want = oracle.multiple_multiexp(curve, pts, sc, chunks)
    assert_same_points(oracle, curve, got, want, f"binned c={c}")
    if chunks == 1 and lines == 1:
        bases_gpu.precompute(14)
        got = engine.multiple_multiexp(w, bases_gpu, sc, 1, 8, True)
        assert w.timings()["scatter_passes"] == 0
        assert_same_points(oracle, curve, got, want, "binned + table")


def test_bn254_batched_4096(engine, oracle, ws):
    """Shape of ag-cuda-ec/benches/multiexp.rs:19-22,56 scaled down: 64 MSMs of 2^12 points."""
    curve, chunks, cl = 0, 64, 4096
    pts, sc = _synth(engine, ws[curve], curve, chunks * cl)
    bases_gpu = engine.upload_multiexp_bases(ws[curve], pts)
    got = engine.multiple_multiexp(ws[curve], bases_gpu, sc, chunks, 8, False)
    want = oracle.multiple_multiexp(curve, pts, sc, chunks)
    assert_same_points(oracle, curve, got, want, "64 x 4096")


@pytest.mark.parametrize("curve", CURVES)
def test_many_small_tasks(engine, oracle, ws, curve):
    """2 lines x 1024 chunks x 16 points: the thread-per-task window combine (n_tasks >= 512)."""
    chunks, cl, lines = 1024, 16, 2
    pts, sc = _synth(engine, ws[curve], curve, chunks * cl * lines)
    sc = sc[: chunks * cl]
    bases_gpu = engine.upload_multiexp_bases(ws[curve], pts)
    got = engine.multiple_multiexp(ws[curve], bases_gpu, sc, chunks, 8, True)
    want = oracle.multiple_multiexp(curve, pts, sc, chunks)
    assert got.shape[0] == lines * chunks
    assert_same_points(oracle, curve, got, want, "2 x 1024 x 16")


def test_bn254_2pow20_bit_exact(engine, oracle, ws):
    """BASELINE.json configs[1]: BN254 G1 MSM 2^20 on one B200, bit-exact vs the CPU multiexp."""
    curve, n = 0, 1 << 20
    pts, sc = _synth(engine, ws[curve], curve, n)
    kern = engine.MultiexpKernel.create([0], curve)
    got = kern.multiexp(engine.Worker(), pts, sc, 0)
    want = oracle.multiexp_cpu(curve, pts, sc)
    assert_same_points(oracle, curve, got, want, "2^20")


def test_linearity_at_full_size(engine, oracle, ws):
    """Size-independent property at BASELINE's full size (2^24 would take the oracle minutes):
    MSM(k, P) over [0, n) == MSM over [0, n/2) + MSM over [n/2, n), and MSM(k, P) with every
    scalar replaced by r - k is the negation."""
    curve, n = 0, 1 << 22
    lib = engine.load_library()
    w = ws[curve]
    pts, sc = _synth(engine, w, curve, n)
    bases_gpu = engine.upload_multiexp_bases(w, pts)
    whole = engine.multiple_multiexp(w, bases_gpu, sc, 1, 8, True)
    halves = engine.multiple_multiexp(w, bases_gpu, sc, 2, 8, True)
    s = oracle.ec_op(curve, 0, halves[0:1].copy(), halves[1:2].copy())
    assert_same_points(oracle, curve, whole, s, "halves")
    r = int.from_bytes(oracle.constant(curve, 5).tobytes(), "little")
    # negate the first 1000 scalars only (python big ints), check the partial identity on a sub-MSM
    m = 1000
    neg = sc[:m].copy()
    for i in range(m):
        k = int.from_bytes(sc[i].tobytes(), "little")
        neg[i] = np.frombuffer(((r - k) % r).to_bytes(32, "little"), dtype=np.uint8)
    sub = engine.upload_multiexp_bases(w, pts[:m])
    a = engine.multiple_multiexp(w, sub, sc[:m], 1, 8, True)
    b = engine.multiple_multiexp(w, sub, neg, 1, 8, True)
    tot = oracle.ec_op(curve, 0, a.copy(), b.copy())
    assert oracle.to_affine(curve, tot)[1].all()
    del lib


def test_context_busy_and_errors(engine, oracle):
    """A context is not re-entrant (CudaError::ContextAlreadyInUse, ag-cuda-proxy/src/context.rs:20-27);
    invalid arguments are reported, not crashed on."""
    lib = engine.load_library()
    w = engine.Workspace(0)
    assert lib.msm_multiple_multiexp(w.handle, None, None, 0, 1, 8, 1, None) != 0
    pts = oracle.gen_points(0, 1, 64)
    bases_gpu = engine.upload_multiexp_bases(w, pts)
    with pytest.raises(engine.CudaError):
        engine.multiple_multiexp(w, bases_gpu, oracle.gen_scalars(0, 1, 128), 1, 8, True)  # L > bases
    flag = ctypes.c_int(1)
    kern = engine.MultiexpKernel.create([0], 0)
    lib.msm_set_abort_flag(kern.workspace.handle, ctypes.addressof(flag))
    with pytest.raises(engine.EcErrorAborted):
        kern.multiexp(engine.Worker(), pts, oracle.gen_scalars(0, 1, 64), 0)
    lib.msm_set_abort_flag(kern.workspace.handle, None)
    kern2 = engine.MultiexpKernel.create_with_abort([0], lambda: True, 0)
    with pytest.raises(engine.EcErrorAborted):
        kern2.multiexp(engine.Worker(), pts, oracle.gen_scalars(0, 1, 64), 0)


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("rounds", [1, 3, 6])
def test_affine_halving_rounds_path(engine, oracle, ws, curve, rounds, monkeypatch):
    """csrc/bucket_affine.cuh (batched-affine bucket accumulation; off by default because it measured slower,
    profiles/r02_affine_rounds.md): forced on, results stay bit-exact -- random inputs with a window table, the
    adversarial set (identity bases, P + P, P - P, equal scalars -> heavy buckets), host scalars in sub-batches."""
    monkeypatch.setenv("MSM_B200_BA_ROUNDS", str(rounds))
    monkeypatch.setenv("MSM_B200_BA_BATCH", "64")
    n = 1 << 17
    pts, sc = _synth(engine, ws[curve], curve, n)
    want = oracle.multiexp_cpu(curve, pts, sc)
    bases = engine.upload_multiexp_bases(ws[curve], pts)
    for call in range(3):  # plain, plain, window table (built by policy on the second call)
        got = engine.multiple_multiexp(ws[curve], bases, sc, 1, 8, True)
        assert_same_points(oracle, curve, got, want, f"affine rounds, call {call}")
    assert bases.table_window() != 0
    bases.free()
    apts, asc = adversarial_inputs(oracle, curve, 1 << 12)
    # identity bases make the reference's CPU path error out; the oracle's batched form skips them like the engine
    abases = engine.upload_multiexp_bases(ws[curve], apts)
    for chunks in (1, 4):
        got = engine.multiple_multiexp(ws[curve], abases, asc, chunks, 8, True)
        assert_same_points(oracle, curve, got, oracle.multiple_multiexp(curve, apts, asc, chunks), f"adversarial, {chunks} chunks")
    abases.free()


@pytest.mark.parametrize("window,neg", [(8, 0), (8, 1), (5, 1)])
def test_reference_cuda_kernel_equals_engine(engine, oracle, ws, window, neg, tmp_path):
    """The reference's OWN kernel (ag-build/cl/multiexp.cl instantiated for BN254 as SourceBuilder does, compiled
    for sm_100a by oracle/build_ref.py --cuda, launched with the geometry of ag-cuda-ec/src/multiexp.rs:27-72) run on
    this GPU against the engine: 2 lines x 16 chunks, equal as group elements -- the parity notion of the reference's
    test_multiexp_batch (ag-cuda-ec/src/multiexp.rs:93-145) with the reference kernel itself as the other side."""
    import os
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_kernel_bn254")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_kernel_bn254 not built (needs /root/reference at build time)")
    curve, L, lines, chunks = 0, 1 << 13, 2, 16
    pts, sc = _synth(engine, ws[curve], curve, L * lines)
    sc = sc[:L]
    fb, fe, fo = (str(tmp_path / x) for x in ("bases.bin", "exps.bin", "out.bin"))
    pts.tofile(fb)
    sc.tofile(fe)
    subprocess.run([exe, fb, fe, str(L), str(chunks), str(window), str(neg), "1", fo], check=True, capture_output=True)
    ref = np.fromfile(fo, dtype=np.uint8).reshape(-1, 3 * FQ[curve])
    bases = engine.upload_multiexp_bases(ws[curve], pts)
    got = engine.multiple_multiexp(ws[curve], bases, sc, chunks, window, bool(neg))
    bases.free()
    assert ref.shape == got.shape == (lines * chunks, 96)
    assert_same_points(oracle, curve, got, ref, f"engine vs reference kernel, window {window}, neg {neg}")


def test_context_busy_under_a_real_race(engine, oracle):
    """Two host threads on ONE workspace: while thread A is inside a long multiple_multiexp call, thread
    B's call on the same context must come back with CudaError::ContextAlreadyInUse (MSM_ERR_BUSY,
    ag-cuda-proxy/src/context.rs:20-27) -- and A's result must be unharmed."""
    import threading

    curve, n = 0, 1 << 20
    w = engine.Workspace(curve)
    pts, sc = _synth(engine, w, curve, n)
    bases = engine.upload_multiexp_bases(w, pts)
    bases.set_table_policy(0)
    small = oracle.gen_scalars(curve, 3, n)
    inside = threading.Event()
    results, busy, other = [], [], []

    def long_call():
        inside.set()
        while len(results) < 6:
            try:
                results.append(engine.multiple_multiexp(w, bases, sc, 1, 8, True))
            except engine.CudaError as e:  # the other thread's call holds the context right now
                (busy if e.name == "ContextAlreadyInUse" else other).append(e)

    a = threading.Thread(target=long_call)
    a.start()
    inside.wait()
    while a.is_alive():
        try:
            engine.multiple_multiexp(w, bases, small, 1, 8, True)
        except engine.CudaError as e:
            (busy if e.name == "ContextAlreadyInUse" else other).append(e)
    a.join()
    assert not other, other
    assert busy, "the second thread never collided with the call in flight"
    want = oracle.multiexp_cpu(curve, pts, sc)
    for r in results:
        assert_same_points(oracle, curve, r, want, "result of the call that held the context")
    bases.free()
    w.close()


def test_abort_mid_call_with_registered_scalars(engine, oracle):
    """Abort flag raised while a pipelined call (pinned host scalars, sub-batch copies on the copy stream) is
    in flight: the call returns OK or Aborted, and in both cases no copy may still be reading the caller's
    buffer -- it is unregistered and overwritten right away, and the next call on the context is correct."""
    import threading
    import time

    curve, n = 0, 1 << 22
    lib = engine.load_library()
    w = engine.Workspace(curve)
    pts, sc = _synth(engine, w, curve, n)
    bases = engine.upload_multiexp_bases(w, pts)
    bases.set_table_policy(0)
    want = oracle.multiexp_cpu(curve, pts, sc)
    flag = ctypes.c_int(0)
    lib.msm_set_abort_flag(w.handle, ctypes.addressof(flag))
    outcomes = set()
    for delay in (0.0, 0.0005, 0.002, 0.005):
        buf = sc.copy()
        assert lib.msm_host_register(buf.ctypes.data, buf.nbytes) == 0
        flag.value = 0
        t = threading.Timer(delay, lambda: setattr(flag, "value", 1))
        t.start()
        try:
            got = engine.multiple_multiexp(w, bases, buf, 1, 8, True)
            assert w.timings()["sub_batches"] > 1
            assert_same_points(oracle, curve, got, want, "call that finished before the abort")
            outcomes.add("ok")
        except engine.CudaError as e:  # ag_cuda_ec has no abort; the C ABI reports MSM_ERR_ABORTED as UnknownError
            outcomes.add("aborted")
            assert "abort" in str(e).lower() or e.name == "UnknownError"
        t.join()
        assert lib.msm_host_unregister(buf.ctypes.data) == 0
        buf[:] = 0xFF  # a copy still in flight would now upload garbage
        time.sleep(0.01)
        flag.value = 0
        again = engine.multiple_multiexp(w, bases, sc, 1, 8, True)
        assert_same_points(oracle, curve, again, want, "call after an aborted one")
    lib.msm_set_abort_flag(w.handle, None)
    bases.free()
    w.close()


def test_too_many_lines_is_an_error_not_a_launch_failure(engine, oracle):
    """n_lines = bases / exponents is unbounded in the reference API; the bucket kernels carry the line in gridDim.y
    (<= 65535): beyond that the call reports MSM_ERR_TOO_LARGE instead of an opaque launch failure, and the context
    stays usable."""
    w = engine.Workspace(0)
    period = oracle.gen_points(0, 3, 64)
    pts = period[np.arange(65536) % 64].copy()
    bases = engine.upload_multiexp_bases(w, pts)
    with pytest.raises(engine.CudaError):
        engine.multiple_multiexp(w, bases, oracle.gen_scalars(0, 3, 1), 1, 8, True)  # 65536 lines of one point
    sc = oracle.gen_scalars(0, 4, 2)
    got = engine.multiple_multiexp(w, bases, sc, 1, 8, True)                           # 32768 lines: fine
    assert_same_points(oracle, 0, got, oracle.multiple_multiexp(0, pts, sc, 1), "32768 lines x 2 points")
    bases.free()
    w.close()


def test_bases_outlive_their_workspace(engine, oracle):
    """Drop order: a DeviceData may be freed after its workspace was closed (msm_b200.h, "Lifetime")."""
    lib = engine.load_library()
    w = engine.Workspace(0)
    bases = engine.upload_multiexp_bases(w, oracle.gen_points(0, 5, 256))
    h = w.handle
    w.close()
    assert lib.msm_multiple_multiexp(h, bases._h, None, 0, 1, 8, 1, None) != 0  # closed context: rejected
    assert bases.size() == 256 * 64
    bases.free()


def test_multi_device_split_and_gather(engine, oracle):
    """MultiexpKernel over every visible GPU: contiguous split ceil(n/devices)
    (ec-gpu-proxy/src/multiexp.rs:329-337), partial points gathered on device 0 over peer copies
    and summed there.  Needs >= 2 GPUs (gpurun --gpus 2); single-GPU boxes skip."""
    lib = engine.load_library()
    ndev = lib.msm_device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    n = (1 << 16) + 123
    for curve in CURVES:
        pts, sc = oracle.gen_points(curve, SEED, n), oracle.gen_scalars(curve, SEED, n)
        kern = engine.MultiexpKernel.create(list(range(ndev)), curve)
        assert kern.num_kernels() == ndev
        want = oracle.multiexp_cpu(curve, pts, sc)
        assert_same_points(oracle, curve, kern.multiexp(engine.Worker(), pts, sc, 0), want, "host bases")
        res = kern.upload_bases(pts)
        assert_same_points(oracle, curve, kern.multiexp_resident(res, sc, 0), want, "resident shards")
        assert lib.msm_bases_precompute(kern.workspace.handle, res, 0) == 0
        assert_same_points(oracle, curve, kern.multiexp_resident(res, sc, 0), want, "resident shards + tables")
        res.free()


def test_concurrent_local_workspaces(engine, oracle):
    """The `_mt` entry points give every host thread its own workspace (construct_workspace!,
    ag-cuda-workspace-macro/src/lib.rs:58-78; ag-cuda-ec/benches/ec_fft.rs:62-110 drives 32 of them at
    once): four threads, each with its own context on the same GPU, run different MSMs concurrently."""
    import threading

    curve = 0
    jobs = []
    for t in range(4):
        n = 3000 + 517 * t
        jobs.append((oracle.gen_points(curve, 100 + t, n), oracle.gen_scalars(curve, 200 + t, n)))
    want = [oracle.multiple_multiexp(curve, p, s, 4) for p, s in jobs]
    got, errs = [None] * 4, []

    def work(t):
        try:
            for _ in range(3):
                bases = engine.upload_multiexp_bases_mt(jobs[t][0], curve)
                got[t] = engine.multiple_multiexp_mt(bases, jobs[t][1], 4, 8, True)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for t in range(4):
        assert_same_points(oracle, curve, got[t], want[t], f"thread {t}")


def test_full_size_default_path_properties(engine, oracle, ws):
    """BASELINE.json configs[2] at its full size, 2^24 BN254 points, through the path bench.py times:
    window table (c = 22), binned sort, host scalars uploaded in 3 - 4 pipelined sub-batches of growing size.  The oracle
    would need minutes at this size, so size-independent properties pin the result:
      * the device-resident call and the pipelined host call agree;
      * the whole MSM equals the sum of 16 chunk MSMs computed without the table (a different window
        size, sort and reduction);
      * a 2^14-point prefix agrees with the oracle bit for bit."""
    curve, n = 0, 1 << 24
    lib = engine.load_library()
    w = ws[curve]
    fq = FQ[curve]
    h = w.handle
    dp, ds, do = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.msm_device_alloc(h, n * 2 * fq, ctypes.byref(dp)) == 0
    assert lib.msm_device_alloc(h, n * 32, ctypes.byref(ds)) == 0
    assert lib.msm_device_alloc(h, 16 * 3 * fq, ctypes.byref(do)) == 0
    try:
        assert lib.msm_synth_points_device(h, SEED, 0, n, dp) == 0
        assert lib.msm_synth_scalars_device(h, SEED, 0, n, ds) == 0
        bh = ctypes.c_void_p()
        assert lib.msm_bases_from_device(h, dp, n, ctypes.byref(bh)) == 0
        # 16 chunks on the plain resident copy
        assert lib.msm_multiple_multiexp_device(h, bh, ds, n, 16, do) == 0
        parts = np.zeros((16, 3 * fq), dtype=np.uint8)
        assert lib.msm_memcpy_d2h(h, parts.ctypes.data, do, parts.nbytes) == 0
        acc = parts[0:1].copy()
        for i in range(1, 16):
            acc = oracle.ec_op(curve, 0, acc, parts[i:i + 1].copy())
        # whole MSM through the table, device-resident
        assert lib.msm_bases_precompute(h, bh, 0) == 0
        assert lib.msm_multiple_multiexp_device(h, bh, ds, n, 1, do) == 0
        t = w.timings()
        assert t["window_bits"] == 22 and t["scatter_passes"] == 0
        whole = np.zeros((1, 3 * fq), dtype=np.uint8)
        assert lib.msm_memcpy_d2h(h, whole.ctypes.data, do, whole.nbytes) == 0
        assert_same_points(oracle, curve, whole, acc, "table + binned sort == sum of 16 plain chunks")
        # the same from host scalars (pipelined sub-batches)
        sc = np.zeros((n, 32), dtype=np.uint8)
        assert lib.msm_memcpy_d2h(h, sc.ctypes.data, ds, sc.nbytes) == 0
        host = np.zeros((1, 3 * fq), dtype=np.uint8)
        assert lib.msm_multiple_multiexp(h, bh, sc.ctypes.data, n, 1, 8, 1, host.ctypes.data) == 0
        first = w.timings()["sub_batches"]
        assert first in (3, 4)  # 4 growing 2-fold; 3 growing 3-fold once an upload rate has been measured on this workspace
        assert_same_points(oracle, curve, host, whole, "pipelined host scalars == device-resident")
        # second host call of the shape: both speeds are known now, the split adapts to them -- same point
        host2 = np.zeros((1, 3 * fq), dtype=np.uint8)
        assert lib.msm_multiple_multiexp(h, bh, sc.ctypes.data, n, 1, 8, 1, host2.ctypes.data) == 0
        assert w.timings()["sub_batches"] in (3, 4)
        assert_same_points(oracle, curve, host2, whole, "adaptive split == device-resident")
        # prefix against the oracle
        m = 1 << 14
        pts = np.zeros((m, 2 * fq), dtype=np.uint8)
        assert lib.msm_memcpy_d2h(h, pts.ctypes.data, dp, pts.nbytes) == 0
        assert (pts == oracle.gen_points(curve, SEED, m)).all() and (sc[:m] == oracle.gen_scalars(curve, SEED, m)).all()
        lib.msm_bases_free(bh)
    finally:
        for d in (dp, ds, do):
            lib.msm_device_free(h, d)
