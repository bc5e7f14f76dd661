#!/usr/bin/env python3
"""Prints the 29-bit-limb constants used by csrc/fp29.cuh (checked by tests/test_constants.py)."""
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
B, N = 29, 9


def limbs(x):
    return [(x >> (B * i)) & ((1 << B) - 1) for i in range(N)]


def fmt(x):
    return ", ".join("0x%08xu" % l for l in limbs(x))


if __name__ == "__main__":
    Rp = 1 << (B * N)
    R = 1 << 256
    print("P        ", fmt(P))
    print("INV      ", hex((-pow(P, -1, 1 << B)) % (1 << B)))
    print("ONE      ", fmt(Rp % P))            # R' mod p
    print("CONV_IN  ", fmt(Rp * Rp * pow(R, -1, P) % P))  # R'^2 / R
    print("CONV_OUT ", fmt(R % P))             # R
