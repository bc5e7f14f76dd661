"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

FQ = {0: 32, 1: 48, 2: 64, 3: 96}  # bytes per coordinate (2, 3: G2 over Fq2)
SEED = 0x0BADC0DE


def affine_of(oracle, curve, jac):
    """Jacobian bytes -> (canonical affine bytes, inf flags) via the oracle."""
    jac = np.ascontiguousarray(jac, dtype=np.uint8).reshape(-1, 3 * FQ[curve])
    return oracle.to_affine(curve, jac, mont_out=False)


def assert_same_points(oracle, curve, got_jac, want_jac, what=""):
    ga, gi = affine_of(oracle, curve, got_jac)
    wa, wi = affine_of(oracle, curve, want_jac)
    assert (gi == wi).all(), f"{what}: infinity flags differ at {np.nonzero(gi != wi)[0][:8]}"
    bad = np.nonzero((ga != wa).any(axis=1))[0]
    assert bad.size == 0, f"{what}: {bad.size} of {len(ga)} results differ, first {bad[:8]}"


def adversarial_inputs(oracle, curve, n, seed=SEED):
    """Edge cases of SURVEY.md section 4: zero / one / r-1 scalars, identity bases, repeated bases
    (same bucket -> doubling branch), P and -P pairs, 99-periodic bases like random_input_by_cycle."""
    fq = FQ[curve]
    period = 99
    base_pts = oracle.gen_points(curve, seed, period)
    pts = base_pts[np.arange(n) % period].copy()
    sc = oracle.gen_scalars(curve, seed, n)
    r = int.from_bytes(oracle.constant(curve, 5).tobytes(), "little")

    def put(i, k):
        sc[i] = np.frombuffer(int(k).to_bytes(32, "little"), dtype=np.uint8)

    if n >= 64:
        put(0, 0)
        put(1, 1)
        put(2, r - 1)
        put(3, 1)
        put(4, (1 << 253) - 1 if r > (1 << 253) else r - 2)
        put(5, 0x8000)       # exactly half of a 16-bit window
        put(6, 0x8001)
        put(7, 0xFFFF)       # carries through
        put(8, (1 << 128) - 1)
        pts[9] = 0           # identity base with a random scalar
        pts[10] = 0
        put(10, 0)
        # P and -P with the same scalar: cancel
        pts[12] = pts[11]
        pts[12, fq:] = oracle.fq_op(curve, 8, pts[11, fq:].copy())
        sc[12] = sc[11]
        # same point, same scalar twice (doubling inside a bucket)
        pts[14] = pts[13]
        sc[14] = sc[13]
        # every scalar of a run identical -> one bucket gets a long run
        for i in range(20, 52):
            sc[i] = sc[19]
    return pts, sc
