"""oracle/pyref.py -- TEST INFRASTRUCTURE ONLY.

Independent pure-Python big-integer model of the MSM path: affine short-Weierstrass
arithmetic over BN254 / BLS12-381 G1, the byte layouts of the reference
(ag-types/src/impls.rs:7-58: bases = {x, y} Montgomery little-endian limbs, identity = (0, 0);
scalars = canonical 32-byte little-endian BigInt<4>; results = Jacobian {x, y, z} Montgomery,
infinity <=> z == 0, ag-build/cl/ec.cl:3-14) and the deterministic synthetic-input generator
(SURVEY.md section 8d).  It shares no code with oracle/msm_oracle.cpp or the CUDA engine and is
what tests/golden/*.json are generated from (tests/golden/make_golden.py).

Only small cases: everything here is plain Python loops.
"""
from __future__ import annotations

MASK64 = (1 << 64) - 1


class CurveParams:
    def __init__(self, name, curve_id, p, r, b, gx, gy, fq_bytes):
        self.name = name
        self.curve_id = curve_id
        self.p = p
        self.r = r
        self.b = b
        self.g = (gx, gy)
        self.fq_bytes = fq_bytes  # 32 (BN254) / 48 (BLS12-381)
        self.R = 1 << (8 * fq_bytes)  # Montgomery radix 2^(32 N)
        self.scalar_bits = r.bit_length()

    # -- Montgomery / byte helpers -------------------------------------------------
    def to_mont(self, x):
        return (x * self.R) % self.p

    def from_mont(self, x):
        return (x * pow(self.R, -1, self.p)) % self.p

    def fq_to_bytes(self, x_mont):
        return int(x_mont).to_bytes(self.fq_bytes, "little")

    def fq_from_bytes(self, b):
        return int.from_bytes(b, "little")

    def affine_to_bytes(self, pt):
        """pt = None (identity -> (0,0), impls.rs:51-57) or (x, y) canonical ints."""
        if pt is None:
            return bytes(2 * self.fq_bytes)
        return self.fq_to_bytes(self.to_mont(pt[0])) + self.fq_to_bytes(self.to_mont(pt[1]))

    def affine_from_bytes(self, b):
        n = self.fq_bytes
        x, y = self.fq_from_bytes(b[:n]), self.fq_from_bytes(b[n : 2 * n])
        if x == 0 and y == 0:
            return None
        return (self.from_mont(x), self.from_mont(y))

    def jacobian_from_bytes(self, b):
        """Jacobian Montgomery {x,y,z} -> affine canonical ints or None."""
        n = self.fq_bytes
        X = self.from_mont(self.fq_from_bytes(b[:n]))
        Y = self.from_mont(self.fq_from_bytes(b[n : 2 * n]))
        Z = self.from_mont(self.fq_from_bytes(b[2 * n : 3 * n]))
        if Z == 0:
            return None
        zi = pow(Z, -1, self.p)
        return (X * zi * zi % self.p, Y * zi * zi * zi % self.p)

    # -- group law (affine, None = infinity) ---------------------------------------
    def on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        return (y * y - x * x * x - self.b) % self.p == 0

    def neg(self, pt):
        if pt is None:
            return None
        return (pt[0], (-pt[1]) % self.p)

    def add(self, a, b):
        if a is None:
            return b
        if b is None:
            return a
        p = self.p
        if a[0] == b[0]:
            if (a[1] + b[1]) % p == 0:
                return None
            lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, p) % p
        else:
            lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, p) % p
        x3 = (lam * lam - a[0] - b[0]) % p
        y3 = (lam * (a[0] - x3) - a[1]) % p
        return (x3, y3)

    def mul(self, k, pt):
        acc = None
        add = pt
        while k:
            if k & 1:
                acc = self.add(acc, add)
            add = self.add(add, add)
            k >>= 1
        return acc

    def msm(self, scalars, points):
        acc = None
        for k, pt in zip(scalars, points):
            acc = self.add(acc, self.mul(k % self.r if k >= self.r else k, pt))
        return acc


BN254 = CurveParams(
    "bn254",
    0,
    0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    3,
    1,
    2,
    32,
)

BLS12_381 = CurveParams(
    "bls12_381",
    1,
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    4,
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
    48,
)

class G2Params(CurveParams):
    """G2 over Fq2 = Fq[u]/(u^2 + 1): coordinates are pairs (c0, c1) of
    canonical integers; bytes are c0 | c1, each in the base field's Montgomery layout (GpuRepr of the
    quadratic extension, ag-types/src/impls.rs:36-46).  `fq_bytes` is the size of one Fq2 coordinate."""

    def __init__(self, name, curve_id, base: CurveParams, b, gx, gy):
        self.name, self.curve_id = name, curve_id
        self.p, self.r, self.b, self.g = base.p, base.r, b, (gx, gy)
        self.base_bytes = base.fq_bytes
        self.fq_bytes = 2 * base.fq_bytes
        self.R = base.R
        self.scalar_bits = base.scalar_bits

    # Fq2 arithmetic on tuples
    def f2add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def f2sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def f2mul(self, a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % self.p, (a[0] * b[1] + a[1] * b[0]) % self.p)

    def f2inv(self, a):
        t = pow(a[0] * a[0] + a[1] * a[1], -1, self.p)
        return (a[0] * t % self.p, -a[1] * t % self.p)

    def fq_to_bytes(self, x_mont):
        return b"".join(int(c).to_bytes(self.base_bytes, "little") for c in x_mont)

    def fq_from_bytes(self, b):
        n = self.base_bytes
        return (int.from_bytes(b[:n], "little"), int.from_bytes(b[n : 2 * n], "little"))

    def to_mont(self, x):
        return tuple(c * self.R % self.p for c in x)

    def from_mont(self, x):
        ri = pow(self.R, -1, self.p)
        return tuple(c * ri % self.p for c in x)

    def affine_from_bytes(self, b):
        n = self.fq_bytes
        x, y = self.fq_from_bytes(b[:n]), self.fq_from_bytes(b[n : 2 * n])
        if x == (0, 0) and y == (0, 0):
            return None
        return (self.from_mont(x), self.from_mont(y))

    def jacobian_from_bytes(self, b):
        n = self.fq_bytes
        X, Y, Z = (self.from_mont(self.fq_from_bytes(b[i * n : (i + 1) * n])) for i in range(3))
        if Z == (0, 0):
            return None
        zi = self.f2inv(Z)
        zi2 = self.f2mul(zi, zi)
        return (self.f2mul(X, zi2), self.f2mul(Y, self.f2mul(zi2, zi)))

    def on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        return self.f2sub(self.f2mul(y, y), self.f2add(self.f2mul(self.f2mul(x, x), x), self.b)) == (0, 0)

    def neg(self, pt):
        return None if pt is None else (pt[0], ((-pt[1][0]) % self.p, (-pt[1][1]) % self.p))

    def add(self, a, b):
        if a is None:
            return b
        if b is None:
            return a
        if a[0] == b[0]:
            if self.f2add(a[1], b[1]) == (0, 0):
                return None
            xx = self.f2mul(a[0], a[0])
            lam = self.f2mul(self.f2add(self.f2add(xx, xx), xx), self.f2inv(self.f2add(a[1], a[1])))
        else:
            lam = self.f2mul(self.f2sub(b[1], a[1]), self.f2inv(self.f2sub(b[0], a[0])))
        x3 = self.f2sub(self.f2sub(self.f2mul(lam, lam), a[0]), b[0])
        y3 = self.f2sub(self.f2mul(lam, self.f2sub(a[0], x3)), a[1])
        return (x3, y3)


def _f2div(p, a, b):
    t = pow(b[0] * b[0] + b[1] * b[1], -1, p)
    bi = (b[0] * t % p, -b[1] * t % p)
    return ((a[0] * bi[0] - a[1] * bi[1]) % p, (a[0] * bi[1] + a[1] * bi[0]) % p)


BN254_G2 = G2Params(
    "bn254_g2", 2, BN254, _f2div(BN254.p, (3, 0), (9, 1)),  # y^2 = x^3 + 3/(9+u)
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
BLS12_381_G2 = G2Params(
    "bls12_381_g2", 3, BLS12_381, (4, 4),  # y^2 = x^3 + 4(1+u)
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)

CURVES = {0: BN254, 1: BLS12_381, 2: BN254_G2, 3: BLS12_381_G2, "bn254": BN254, "bls12_381": BLS12_381,
          "bn254_g2": BN254_G2, "bls12_381_g2": BLS12_381_G2}

# Public known answers (canonical affine): EIP-196 / py_ecc alt_bn128 multiples of G.
BN254_2G = (
    0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
    0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4,
)
BN254_3G = (
    0x0769BF9AC56BEA3FF40232BCB1B6BD159315D84715B8E679F2D355961915ABF0,
    0x2AB799BEE0489429554FDB7C8D086475319E63B40B9C5B57CDF1FF3DD9FE2261,
)


# -- deterministic synthetic inputs (must match msm_oracle.cpp and the CUDA generator) --------
def splitmix64(seed, idx):
    z = (seed + (idx + 1) * 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def gen_scalar(cv: CurveParams, seed, i):
    bits = cv.scalar_bits
    top_mask = (1 << (bits % 64)) - 1 if bits % 64 else MASK64
    attempt = 0
    while True:
        s = (seed + attempt * 0xD1B54A32D192ED03) & MASK64
        w = [splitmix64(s, 4 * i + j) for j in range(4)]
        w[3] &= top_mask
        k = w[0] | (w[1] << 64) | (w[2] << 128) | (w[3] << 192)
        if k < cv.r:
            return k
        attempt += 1


def gen_scalars(cv, seed, start, n):
    return [gen_scalar(cv, seed, start + i) for i in range(n)]


def gen_point_scalars(seed):
    a = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5, 0)
    b = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5, 1) | 1
    return a, b


def gen_points(cv, seed, start, n):
    """P_i = (a + i*b) * G, canonical affine."""
    a, b = gen_point_scalars(seed)
    d = cv.mul(b, cv.g)
    cur = cv.mul(a + start * b, cv.g)
    out = []
    for _ in range(n):
        out.append(cur)
        cur = cv.add(cur, d)
    return out


def scalars_to_bytes(scalars):
    return b"".join(int(k).to_bytes(32, "little") for k in scalars)


def points_to_bytes(cv, pts):
    return b"".join(cv.affine_to_bytes(p) for p in pts)
