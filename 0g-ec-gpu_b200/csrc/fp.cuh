// fp.cuh -- prime-field arithmetic on N 32-bit limbs in Montgomery form (R = 2^(32 N)).
//
// Replaces the generated field layer of the reference: FIELD_add/sub/double/mul/sqr
// (ag-build/cl/field.cl:14-69, 85-263, 313-325) and the per-field constants emitted by
// ag-build/src/source/template.rs:35-71.  Constants are fixed templates for BN254 Fq and
// BLS12-381 Fq (values cross-checked against oracle/pyref.py in tests/test_constants.py).
//
// All values are kept fully reduced in [0, p).  Layout: little-endian uint32 limbs, identical to
// the reference's FIELD struct (ag-build/src/source/template.rs:52).
#pragma once
#include "ptx.cuh"

namespace msm {

// ---------------------------------------------------------------------------------------------
// Field parameter packs.  P(i) / ONE(i) / R2(i) fold to immediates after unrolling.
// ---------------------------------------------------------------------------------------------
struct Bn254Fq {
  static constexpr int N = 8;
  static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32 (limb.rs:65-72)
  static MSM_HD constexpr uint32_t P(int i) {
    constexpr uint32_t t[N] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return t[i];
  }
  static MSM_HD constexpr uint32_t ONE(int i) {  // R mod p
    constexpr uint32_t t[N] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  static MSM_HD constexpr uint32_t R2(int i) {  // R^2 mod p
    constexpr uint32_t t[N] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return t[i];
  }
  static constexpr uint32_t CURVE_B = 3;  // y^2 = x^3 + 3
};

struct Bls381Fq {
  static constexpr int N = 12;
  static constexpr uint32_t INV = 0xfffcfffdu;
  static MSM_HD constexpr uint32_t P(int i) {
    constexpr uint32_t t[N] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu,
                               0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u,
                               0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return t[i];
  }
  static MSM_HD constexpr uint32_t ONE(int i) {
    constexpr uint32_t t[N] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu,
                               0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u,
                               0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
    return t[i];
  }
  static MSM_HD constexpr uint32_t R2(int i) {
    constexpr uint32_t t[N] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u,
                               0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u,
                               0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
    return t[i];
  }
  static constexpr uint32_t CURVE_B = 4;  // y^2 = x^3 + 4
};

// Scalar fields Fr (only Montgomery -> canonical conversion of exponents runs on them:
// PrimeFieldRepr::to_bigint, ag-types/src/impls.rs:7-18, moved onto the device).
struct Bn254Fr {
  static constexpr int N = 8;
  static constexpr uint32_t INV = 0xefffffffu;
  static MSM_HD constexpr uint32_t P(int i) {
    constexpr uint32_t t[N] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return t[i];
  }
  static MSM_HD constexpr uint32_t ONE(int i) {
    constexpr uint32_t t[N] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  static MSM_HD constexpr uint32_t R2(int i) {
    constexpr uint32_t t[N] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return t[i];
  }
};
struct Bls381Fr {  // 255-bit modulus: p < 2^(32N-1) still satisfies the bound of mont_row
  static constexpr int N = 8;
  static constexpr uint32_t INV = 0xffffffffu;
  static MSM_HD constexpr uint32_t P(int i) {
    constexpr uint32_t t[N] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                               0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    return t[i];
  }
  static MSM_HD constexpr uint32_t ONE(int i) {
    constexpr uint32_t t[N] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                               0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
    return t[i];
  }
  static MSM_HD constexpr uint32_t R2(int i) {
    constexpr uint32_t t[N] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                               0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
    return t[i];
  }
};

// ---------------------------------------------------------------------------------------------
template <class P> struct Fp {
  static constexpr int N = P::N;
  uint32_t v[N];
};

template <class P> MSM_HD Fp<P> fp_zero() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.v[i] = 0;
  return r;
}
template <class P> MSM_HD Fp<P> fp_one() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.v[i] = P::ONE(i);
  return r;
}
template <class P> MSM_HD bool fp_is_zero(const Fp<P>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) o |= a.v[i];
  return o == 0;
}
template <class P> MSM_HD bool fp_eq(const Fp<P>& a, const Fp<P>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

// r = (r >= p) ? r - p : r      (r < 2p on entry)
template <class P> MSM_HD void fp_csub_p(uint32_t* r) {
  constexpr int N = P::N;
  uint32_t t[N];
  t[0] = sub_cc(r[0], P::P(0));
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = subc_cc(r[i], P::P(i));
  uint32_t borrow = subc(0u, 0u);  // 0xffffffff when r < p
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

// limb i of K*p (K = 1 or 2), compile-time
template <class P, int K> MSM_HD constexpr uint32_t fp_kp(int i) {
  return K == 1 ? P::P(i) : (uint32_t)((P::P(i) << 1) | (i ? (P::P(i - 1) >> 31) : 0u));
}
// r = (r >= K p) ? r - K p : r
template <class P, int K> MSM_HD void fp_csub_kp(uint32_t* r) {
  constexpr int N = P::N;
  uint32_t t[N];
  t[0] = sub_cc(r[0], fp_kp<P, K>(0));
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = subc_cc(r[i], fp_kp<P, K>(i));
  uint32_t borrow = subc(0u, 0u);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

template <class P> MSM_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);  // p < 2^(32N-2): no carry out
  fp_csub_p<P>(r.v);
  return r;
}

template <class P> MSM_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t mask = subc(0u, 0u);  // 0xffffffff when a < b
  r.v[0] = add_cc(r.v[0], P::P(0) & mask);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], P::P(i) & mask);
  r.v[N - 1] = addc(r.v[N - 1], P::P(N - 1) & mask);
  return r;
}

template <class P> MSM_HD Fp<P> fp_dbl(const Fp<P>& a) { return fp_add<P>(a, a); }

template <class P> MSM_HD Fp<P> fp_neg(const Fp<P>& a) {
  constexpr int N = P::N;
  // p - a, mapped to 0 when a == 0
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < N; i++) nz |= a.v[i];
  uint32_t mask = nz ? 0xffffffffu : 0u;
  Fp<P> r;
  r.v[0] = sub_cc(P::P(0) & mask, a.v[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(P::P(i) & mask, a.v[i]);
  r.v[N - 1] = subc(P::P(N - 1) & mask, a.v[N - 1]);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Montgomery product, coarsely-integrated operand scanning with two accumulators that hold the
// even-aligned (E) and odd-aligned (O) 64-bit columns, so that every 32x32->64 product is one
// wide multiply-add on a 64-bit-aligned register pair and the carry chain of a row never has to
// ripple across the whole accumulator.  Per row i:   T += a*b[i];  m = T*INV mod 2^32;
// T += m*p;  T >>= 32.  The shift swaps the roles of E and O (the old E, shifted by 64 bits,
// becomes the new O; its orphan limb E[1] is folded into the new E[0] with the carry entering
// the next row's O chain).  2N^2 + N multiplier-pipe operations (N=8: 136).
//
// Invariant: T = sum E[k] 2^(32k) + 2^32 sum O[k] 2^(32k) < 2^(32(N+1)) at every point (because
// p < 2^(32N-2)), hence neither chain ever carries out of O[N-1] and E's carry-out is added to
// O[N-1] (same weight 2^(32N)).
// ---------------------------------------------------------------------------------------------
template <class P, bool FIRST>
MSM_HD void mont_row(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
  constexpr int N = P::N;
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      mul_wide(E[j], E[j + 1], a[j], bi);
      mul_wide(O[j], O[j + 1], a[j + 1], bi);
    }
  } else {
    // E is last row's O; O is last row's E (its limb 0 is zero, limb 1 is the orphan).
    E[0] = add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) madc_wide_cc3(O[j], O[j + 1], a[j + 1], bi, O[j + 2], O[j + 3]);
    madc_wide_0(O[N - 2], O[N - 1], a[N - 1], bi);
    mad_wide_cc(E[0], E[1], a[0], bi);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], a[j], bi);
    O[N - 1] = addc(O[N - 1], 0u);
  }
  const uint32_t m = mul_lo(E[0], P::INV);
  mad_wide_cc(O[0], O[1], P::P(1), m);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(O[j], O[j + 1], P::P(j + 1), m);
  mad_wide_cc(E[0], E[1], P::P(0), m);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], P::P(j), m);
  O[N - 1] = addc(O[N - 1], 0u);
}

// a*b/R without the final conditional subtraction: result < a*b/R + p (< 2p for a, b < 2p)
template <class P> MSM_HD Fp<P> fp_mul_nored(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  static_assert(N % 2 == 0, "even limb count required");
  uint32_t X[N], Y[N];
  mont_row<P, true>(X, Y, a.v, b.v[0]);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_row<P, false>(Y, X, a.v, b.v[i]);
    if (i + 1 < N) mont_row<P, false>(X, Y, a.v, b.v[i + 1]);
  }
  // last row had E = Y, O = X:  result = (E >> 32) + O
  Fp<P> r;
  r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
  r.v[N - 1] = addc(X[N - 1], 0u);
  return r;
}
template <class P> MSM_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r = fp_mul_nored<P>(a, b);
  fp_csub_p<P>(r.v);
  return r;
}

// (a*b + c*d)/R with ONE Montgomery reduction: each row adds a*b[i] and c*d[i] before the
// quotient digit is taken.  Saves N^2+N of the 4N^2+2N multiplies of two separate products.
// Bound: T < 5 p 2^32 + 2p < 2^(32(N+1)) for a, b, c, d < 2p (p < 0.19 * 2^(32N)), so the carry
// structure of mont_row is unchanged.  Result < (ab + cd)/R + p, not reduced.
template <class P, bool FIRST>
MSM_HD void mont_row2(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, const uint32_t* c, uint32_t di) {
  constexpr int N = P::N;
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      mul_wide(E[j], E[j + 1], a[j], bi);
      mul_wide(O[j], O[j + 1], a[j + 1], bi);
    }
  } else {
    E[0] = add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) madc_wide_cc3(O[j], O[j + 1], a[j + 1], bi, O[j + 2], O[j + 3]);
    madc_wide_0(O[N - 2], O[N - 1], a[N - 1], bi);
    mad_wide_cc(E[0], E[1], a[0], bi);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], a[j], bi);
    O[N - 1] = addc(O[N - 1], 0u);
  }
  // second product of the row
  mad_wide_cc(O[0], O[1], c[1], di);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(O[j], O[j + 1], c[j + 1], di);
  mad_wide_cc(E[0], E[1], c[0], di);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], c[j], di);
  O[N - 1] = addc(O[N - 1], 0u);
  const uint32_t m = mul_lo(E[0], P::INV);
  mad_wide_cc(O[0], O[1], P::P(1), m);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(O[j], O[j + 1], P::P(j + 1), m);
  mad_wide_cc(E[0], E[1], P::P(0), m);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], P::P(j), m);
  O[N - 1] = addc(O[N - 1], 0u);
}
template <class P> MSM_HD Fp<P> fp_mul2_nored(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
  constexpr int N = P::N;
  uint32_t X[N], Y[N];
  mont_row2<P, true>(X, Y, a.v, b.v[0], c.v, d.v[0]);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_row2<P, false>(Y, X, a.v, b.v[i], c.v, d.v[i]);
    if (i + 1 < N) mont_row2<P, false>(X, Y, a.v, b.v[i + 1], c.v, d.v[i + 1]);
  }
  Fp<P> r;
  r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
  r.v[N - 1] = addc(X[N - 1], 0u);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Dedicated squaring (the reference has none: FIELD_sqr = FIELD_mul(a, a), ag-build/cl/field.cl:313-315).
//
//   a^2 = sum_i a_i 2^(32 i) * V_i,   V_i = a_i 2^(32 i) + 2 * sum_{j>i} a_j 2^(32 j)
//
// so row i of the operand-scanning product above multiplies by a_i a vector V_i whose limbs below i are zero:
// limb i is a_i, limb i+1 is d_(i+1) with bit 0 cleared, limb j > i+1 is d_j, where d = 2a (one funnel shift
// per limb; a < 2p < 2^(32N-1), so d fits N limbs; bit 0 of d_(i+1) is the top bit of a_i, which belongs to
// 2 a_i and not to the tail).  In the even/odd row form only the skipped products of the EVEN accumulator are
// free: the odd accumulator's chain also carries the 64-bit shift of the row (madc_wide_cc3 reads [j+2]) and
// the carry of the orphan limb, so a skipped product there still costs two carry adds -- the pipe time of the
// IMAD.WIDE.X it replaces.  (Moving the orphan's carry elsewhere does not work: it has the weight of the low
// limb of exactly that chain's first pair.)  Products executed: N^2 - sum_i ceil(i/2) = 48 of 64 for N = 8,
// 108 of 144 for N = 12, plus the N^2 + N of the reduction, against 15 / 23 shift-and-mask instructions for d:
// 6 - 8 % of a squaring's pipe time (measured: DESIGN.md section 3).  A full N(N+1)/2 squaring needs the
// separated (product, then reduction) form, whose 2N-limb intermediate costs more carry adds on this chip,
// where IADD3.X shares the multiplier's datapath, than the 12 / 30 further products it saves.
// The doubled vector is below 4p, so a row adds less than 5p 2^32 to T: the accumulator invariant needs
// 5p < 2^(32N) (true for BN254 Fq / Fr and BLS12-381 Fq; BLS12-381 Fr, p = 0.45 * 2^256, keeps the product).
// ---------------------------------------------------------------------------------------------
template <class P> MSM_HD constexpr bool fp_has_fast_sqr() { return P::P(P::N - 1) < 0x33333333u; }

// Row I >= 1 of the squaring: mont_row<P, false> with v[j] = 0 for j < I.
template <class P, int I>
MSM_HD void mont_row_sq(uint32_t* E, uint32_t* O, const uint32_t* v, uint32_t bi) {
  constexpr int N = P::N;
  constexpr int JE = I + (I & 1);  // first even limb >= I
  E[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 0; j < N - 2; j += 2) {
    if (j + 1 < I) {  // product skipped: shift and carry only
      O[j] = addc_cc(O[j + 2], 0u);
      O[j + 1] = addc_cc(O[j + 3], 0u);
    } else {
      madc_wide_cc3(O[j], O[j + 1], v[j + 1], bi, O[j + 2], O[j + 3]);
    }
  }
  madc_wide_0(O[N - 2], O[N - 1], v[N - 1], bi);
  if (JE <= N - 2) {
    mad_wide_cc(E[JE], E[JE + 1], v[JE], bi);
#pragma unroll
    for (int j = JE + 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], v[j], bi);
    O[N - 1] = addc(O[N - 1], 0u);
  }
  const uint32_t m = mul_lo(E[0], P::INV);
  mad_wide_cc(O[0], O[1], P::P(1), m);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(O[j], O[j + 1], P::P(j + 1), m);
  mad_wide_cc(E[0], E[1], P::P(0), m);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], P::P(j), m);
  O[N - 1] = addc(O[N - 1], 0u);
}

template <class P, int I> struct SqRows {
  // rows I, I+1: (E, O) = (Y, X) then (X, Y), as in fp_mul_nored
  static MSM_HD void run(uint32_t* X, uint32_t* Y, const uint32_t* a, const uint32_t* d) {
    constexpr int N = P::N;
    uint32_t v[N];
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = j < I ? 0u : (j == I ? a[I] : (j == I + 1 ? (d[j] & ~1u) : d[j]));
    mont_row_sq<P, I>(Y, X, v, a[I]);
    if (I + 1 < N) {
      uint32_t w[N];
#pragma unroll
      for (int j = 0; j < N; j++) w[j] = j < I + 1 ? 0u : (j == I + 1 ? a[I + 1] : (j == I + 2 ? (d[j] & ~1u) : d[j]));
      mont_row_sq<P, I + 1>(X, Y, w, a[I + 1]);
    }
    SqRows<P, I + 2>::run(X, Y, a, d);
  }
};
template <class P> struct SqRows<P, P::N + 1> {
  static MSM_HD void run(uint32_t*, uint32_t*, const uint32_t*, const uint32_t*) {}
};

// a^2 / R without the final conditional subtraction; a < 2p; result < a^2/R + p (< 2p)
template <class P> MSM_HD Fp<P> fp_sqr_nored(const Fp<P>& a) {
  constexpr int N = P::N;
  static_assert(N % 2 == 0, "even limb count required");
  if (!fp_has_fast_sqr<P>()) return fp_mul_nored<P>(a, a);
  uint32_t d[N];
  d[0] = a.v[0] << 1;
#pragma unroll
  for (int j = 1; j < N; j++) d[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
  uint32_t X[N], Y[N], v0[N];
#pragma unroll
  for (int j = 0; j < N; j++) v0[j] = j == 0 ? a.v[0] : (j == 1 ? (d[1] & ~1u) : d[j]);
  mont_row<P, true>(X, Y, v0, a.v[0]);
  SqRows<P, 1>::run(X, Y, a.v, d);
  // last row had E = Y, O = X:  result = (E >> 32) + O
  Fp<P> r;
  r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
  r.v[N - 1] = addc(X[N - 1], 0u);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Carry-save variant of the same product.  On B200 the carry-in form IMAD.WIDE.U32.X holds the
// multiplier pipe for twice as long as a carry-free IMAD.WIDE.U32 (tools/imad_peak.cu,
// profiles/r01_imad_peak_v2.json), and a field product is ~100 of them.  Here NO multiply takes a
// carry in: each wide multiply-add's carry OUT is counted (one IADD3.X on the otherwise idle ALU
// pipe) in a small counter attached to the limb two positions up, and the counters are folded in
// once per product.  Same even/odd accumulators and the same invariant as mont_row above; cE[k] /
// cO[k] = pending carries into E[k] / O[k].  Counters never exceed 4N.
// ---------------------------------------------------------------------------------------------
template <class P, bool FIRST>
MSM_HD void mont_row_cs(uint32_t* E, uint32_t* O, uint32_t* cE, uint32_t* cO, const uint32_t* a, uint32_t bi) {
  constexpr int N = P::N;
  // on entry (non-first): E/cE are last row's O/cO; O/cO are last row's E/cE, still unshifted
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      mul_wide(E[j], E[j + 1], a[j], bi);
      mul_wide(O[j], O[j + 1], a[j + 1], bi);
    }
#pragma unroll
    for (int j = 0; j < N; j++) cE[j] = cO[j] = 0;
  } else {
    // orphan limb O[1] (and the carries still pending into it) joins the new E[0]; their carries
    // are pending into E[1]
    uint32_t c1 = cE[1];
    add_cs(E[0], c1, cE[0]);  // carries that were pending into this limb while it was O[0]
    add_cs(E[0], c1, O[1]);
    add_cs(E[0], c1, cO[1]);
    cE[0] = 0;
    cE[1] = c1;
    // shift the old E (now O) down by two limbs while adding the odd products: new O[j] = old
    // O[j+2] and new cO[j] = old cO[j+2] (nothing ever carries into limb 0); the product
    // a[j+1]*bi lands on the pair (j, j+1) and its carry on limb j+2
    uint32_t nO[N], ncO[N];
#pragma unroll
    for (int j = 0; j < N; j++) ncO[j] = j + 2 < N ? cO[j + 2] : 0;
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) mad_wide_cs3(nO[j], nO[j + 1], ncO[j + 2], a[j + 1], bi, O[j + 2], O[j + 3]);
    mul_wide(nO[N - 2], nO[N - 1], a[N - 1], bi);
#pragma unroll
    for (int j = 0; j < N; j++) {
      O[j] = nO[j];
      cO[j] = ncO[j];
    }
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) mad_wide_cs(E[j], E[j + 1], cE[j + 2], a[j], bi);
    mad_wide_cs(E[N - 2], E[N - 1], cO[N - 1], a[N - 2], bi);  // weight 2^(32N) = limb O[N-1]
  }
  const uint32_t m = mul_lo(E[0], P::INV);
#pragma unroll
  for (int j = 0; j < N - 2; j += 2) mad_wide_cs(O[j], O[j + 1], cO[j + 2], P::P(j + 1), m);
  {
    uint32_t never = 0;  // the top pair cannot overflow (T < 2^(32(N+1)))
    mad_wide_cs(O[N - 2], O[N - 1], never, P::P(N - 1), m);
  }
#pragma unroll
  for (int j = 0; j < N - 2; j += 2) mad_wide_cs(E[j], E[j + 1], cE[j + 2], P::P(j), m);
  mad_wide_cs(E[N - 2], E[N - 1], cO[N - 1], P::P(N - 2), m);
}

template <class P> MSM_HD Fp<P> fp_mul_cs(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  uint32_t X[N], Y[N], cX[N], cY[N];
  mont_row_cs<P, true>(X, Y, cX, cY, a.v, b.v[0]);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_row_cs<P, false>(Y, X, cY, cX, a.v, b.v[i]);
    if (i + 1 < N) mont_row_cs<P, false>(X, Y, cX, cY, a.v, b.v[i + 1]);
  }
  // last row had E = Y, O = X:  result = (E >> 32) + O + pending carries (cE[k+1] + cO[k] into limb k)
  Fp<P> r;
  r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
  r.v[N - 1] = addc(X[N - 1], 0u);
  uint32_t cnt[N];
#pragma unroll
  for (int k = 0; k < N; k++) cnt[k] = cX[k] + (k + 1 < N ? cY[k + 1] : 0u);
  r.v[0] = add_cc(r.v[0], cnt[0]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(r.v[k], cnt[k]);
  r.v[N - 1] = addc(r.v[N - 1], cnt[N - 1]);
  fp_csub_p<P>(r.v);
  return r;
}

template <class P> MSM_HD Fp<P> fp_sqr(const Fp<P>& a) {
  Fp<P> r = fp_sqr_nored<P>(a);
  fp_csub_p<P>(r.v);
  return r;
}

// Montgomery <-> canonical
template <class P> MSM_HD Fp<P> fp_to_mont(const Fp<P>& a) {
  Fp<P> r2;
#pragma unroll
  for (int i = 0; i < P::N; i++) r2.v[i] = P::R2(i);
  return fp_mul<P>(a, r2);
}
template <class P> MSM_HD Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> o = fp_zero<P>();
  o.v[0] = 1;
  return fp_mul<P>(a, o);
}

// a^(p-2); a != 0.  Only used once per result / per synthetic-input batch, never in the hot loop.
template <class P> MSM_COLD Fp<P> fp_inv(const Fp<P>& a) {
  constexpr int N = P::N;
  uint32_t e[N];
  e[0] = sub_cc(P::P(0), 2u);
#pragma unroll
  for (int i = 1; i < N; i++) e[i] = subc_cc(P::P(i), 0u);
  Fp<P> acc = fp_one<P>();
  Fp<P> base = a;
  for (int i = 0; i < 32 * N; i++) {
    if ((e[i >> 5] >> (i & 31)) & 1) acc = fp_mul<P>(acc, base);
    base = fp_sqr<P>(base);
  }
  return acc;
}

}  // namespace msm

// ---------------------------------------------------------------------------------------------
// FieldSat<P>: the saturated 32-bit-limb field above behind the static interface the curve and
// MSM templates use (see fp29.cuh for the lazily reduced 29-bit alternative).  Values are always
// canonical here, so the lazy-reduction hooks (K multiples of p, norm) are no-ops.
// ---------------------------------------------------------------------------------------------
namespace msm {

// CS = true: products use the carry-save form fp_mul_cs.
template <class P, bool CS = false> struct FieldSat {
  static constexpr int N = P::N;
  static constexpr int API_WORDS = P::N;     // words per element at the API boundary
  static constexpr int PACKED_WORDS = P::N;  // words per resident-base coordinate
  using Elem = Fp<P>;

  static MSM_HD Elem zero() { return fp_zero<P>(); }
  static MSM_HD Elem one() { return fp_one<P>(); }
  static MSM_HD bool is_zero_limbs(const Elem& a) { return fp_is_zero<P>(a); }
  static MSM_HD Elem add(const Elem& a, const Elem& b) { return fp_add<P>(a, b); }
  template <int K, int LM> static MSM_HD Elem sub(const Elem& a, const Elem& b) { return fp_sub<P>(a, b); }
  template <int K, int LM> static MSM_HD Elem neg(const Elem& a) { return fp_neg<P>(a); }
  static MSM_HD Elem norm(const Elem& a) { return a; }
  template <int LO, int HI> static MSM_HD bool is_multiple_of_p(const Elem& a) { return fp_is_zero<P>(a); }
  static MSM_HD Elem mul(const Elem& a, const Elem& b) { return CS ? fp_mul_cs<P>(a, b) : fp_mul<P>(a, b); }
  static MSM_HD Elem sqr(const Elem& a) { return mul(a, a); }
  // a*b - c*d with one reduction: (a*b + (p - c)*d)/R < 2p^2/R + p < 2p
  static MSM_HD Elem mul_sub(const Elem& a, const Elem& b, const Elem& c, const Elem& d) {
    Elem r = fp_mul2_nored<P>(a, b, fp_neg<P>(c), d);
    fp_csub_p<P>(r.v);
    return r;
  }
  static MSM_HD Elem inv(const Elem& a) { return fp_inv<P>(a); }

  // resident-base coordinate <-> registers (same layout as the API: nothing to do)
  static MSM_HD Elem unpack(const uint32_t* w) {
    Elem r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = w[i];
    return r;
  }
  static MSM_HD void api_to_packed(const uint32_t* api, uint32_t* packed) {
#pragma unroll
    for (int i = 0; i < N; i++) packed[i] = api[i];
  }
  static MSM_HD void to_packed(const Elem& a, uint32_t* packed) {
#pragma unroll
    for (int i = 0; i < N; i++) packed[i] = a.v[i];
  }
  static MSM_HD Elem from_api(const uint32_t* w) { return unpack(w); }
  static MSM_HD void to_api(const Elem& a, uint32_t* w) {
#pragma unroll
    for (int i = 0; i < N; i++) w[i] = a.v[i];
  }
  // canonical integer (non-Montgomery) in API word layout
  static MSM_HD void to_canonical_words(const Elem& a, uint32_t* w) {
    Elem c = fp_from_mont<P>(a);
#pragma unroll
    for (int i = 0; i < N; i++) w[i] = c.v[i];
  }
};

// ---------------------------------------------------------------------------------------------
// FieldSatLazy<P>: the same 32-bit-limb arithmetic with values kept in [0, 2p) instead of [0, p).
// A Montgomery product of two values < 2p is < 4p^2/R + p < 2p (p < 0.19 R for both curves), so the
// conditional subtraction after every product -- 17 of the ~190 instructions, none of which
// overlaps with the multiplier pipe on B200 -- disappears.  add/sub fold back into [0, 2p).
// Zero tests compare against 0 and p; the infinity flag is still "all limbs zero" because a finite
// point's ZZ is a product of non-zero residues and is never 0 or p.
// ---------------------------------------------------------------------------------------------
template <class P> struct FieldSatLazy {
  static constexpr int N = P::N;
  static constexpr int API_WORDS = P::N;
  static constexpr int PACKED_WORDS = P::N;
  using Elem = Fp<P>;

  static MSM_HD Elem zero() { return fp_zero<P>(); }
  static MSM_HD Elem one() { return fp_one<P>(); }
  static MSM_HD bool is_zero_limbs(const Elem& a) { return fp_is_zero<P>(a); }
  static MSM_HD Elem add(const Elem& a, const Elem& b) {
    Elem r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
    r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);  // 4p < 2^(32N)
    fp_csub_kp<P, 2>(r.v);
    return r;
  }
  // (Round 2 tried the correction as one asm block with a predicated add chain and the limbs of 2p as immediates --
  // 2N + 2 instructions instead of 3N + 1: BN254 unchanged (29.36 against 29.41 ms of accumulation at 2^24),
  // BLS12-381 4.5 % slower (19.66 against 18.82 ms at 2^22): the block is a scheduling barrier for ptxas.  Job r2_run18.)
  template <int K, int LM> static MSM_HD Elem sub(const Elem& a, const Elem& b) {
    Elem r;
    r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
    const uint32_t mask = subc(0u, 0u);
    r.v[0] = add_cc(r.v[0], fp_kp<P, 2>(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], fp_kp<P, 2>(i) & mask);
    r.v[N - 1] = addc(r.v[N - 1], fp_kp<P, 2>(N - 1) & mask);
    return r;
  }
  template <int K, int LM> static MSM_HD Elem neg(const Elem& a) {
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < N; i++) nz |= a.v[i];
    const uint32_t mask = nz ? 0xffffffffu : 0u;
    Elem r;
    r.v[0] = sub_cc(fp_kp<P, 2>(0) & mask, a.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(fp_kp<P, 2>(i) & mask, a.v[i]);
    r.v[N - 1] = subc(fp_kp<P, 2>(N - 1) & mask, a.v[N - 1]);
    return r;
  }
  static MSM_HD Elem norm(const Elem& a) { return a; }
  template <int LO, int HI> static MSM_HD bool is_multiple_of_p(const Elem& a) {
    uint32_t z = 0, e = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      z |= a.v[i];
      e |= a.v[i] ^ P::P(i);
    }
    return z == 0 || e == 0;
  }
  static MSM_HD Elem mul(const Elem& a, const Elem& b) { return fp_mul_nored<P>(a, b); }
  static MSM_HD Elem sqr(const Elem& a) { return fp_sqr_nored<P>(a); }
  // a*b - c*d: (a*b + (2p - c)*d)/R < 8p^2/R + p < 4p -> one fold back below 2p
  static MSM_HD Elem mul_sub(const Elem& a, const Elem& b, const Elem& c, const Elem& d) {
    Elem r = fp_mul2_nored<P>(a, b, neg<2, 1>(c), d);
    fp_csub_kp<P, 2>(r.v);
    return r;
  }
  static MSM_HD Elem canonical(const Elem& a) {
    Elem r = a;
    fp_csub_p<P>(r.v);
    return r;
  }
  static MSM_COLD Elem inv(const Elem& a) { return fp_inv<P>(canonical(a)); }

  static MSM_HD Elem unpack(const uint32_t* w) {
    Elem r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = w[i];
    return r;
  }
  static MSM_HD void api_to_packed(const uint32_t* api, uint32_t* packed) {
#pragma unroll
    for (int i = 0; i < N; i++) packed[i] = api[i];
  }
  static MSM_HD void to_packed(const Elem& a, uint32_t* packed) { to_api(a, packed); }
  static MSM_HD Elem from_api(const uint32_t* w) { return unpack(w); }
  static MSM_HD void to_api(const Elem& a, uint32_t* w) {
    const Elem c = canonical(a);
#pragma unroll
    for (int i = 0; i < N; i++) w[i] = c.v[i];
  }
  static MSM_HD void to_canonical_words(const Elem& a, uint32_t* w) {
    Elem c = fp_from_mont<P>(canonical(a));
#pragma unroll
    for (int i = 0; i < N; i++) w[i] = c.v[i];
  }
};

}  // namespace msm
