#!/usr/bin/env python3
"""Exactness check of tools/dfma_mul samples: r == a*b*2^(-52N) mod p and r < a*b/R + p."""
import sys
P = {"bn254": 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
     "bls12_381": 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab}
bad = n = 0
for line in open(sys.argv[1]):
    f = line.split()
    name, N = f[0], int(f[1])
    v = [int(x, 16) for x in f[2:]]
    val = lambda l: sum(x << (52 * i) for i, x in enumerate(l))
    a, b, r = val(v[:N]), val(v[N:2 * N]), val(v[2 * N:3 * N])
    p, R = P[name], 1 << (52 * N)
    ok = (r * R - a * b) % p == 0 and r < a * b // R + p + 1 and all(x < (1 << 52) for x in v[2 * N:3 * N])
    bad += not ok
    n += 1
print("dfma samples checked: %d, bad: %d" % (n, bad))
sys.exit(1 if bad or n == 0 else 0)
