// ec.cuh -- short-Weierstrass a = 0 group law for the MSM buckets, generic over the field class
// (FieldSat<P> of fp.cuh: canonical 32-bit limbs; FieldU29<U> of fp29.cuh: lazy 29-bit limbs).
//
// Replaces POINT_add_mixed / POINT_add / POINT_double of ag-build/cl/ec.cl:17-120.  Buckets are
// kept in extended Jacobian ("XYZZ") coordinates, x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2 (EFD
// "shortw/xyzz": madd-2008-s = 8M+2S, add-2008-s = 12M+2S, dbl-2008-s-1, mdbl-2008-s) instead of
// the reference's Jacobian madd-2007-bl (7M+4S, executed as 11 multiplies).  Results leave the
// engine as Jacobian {x,y,z} Montgomery with infinity <=> z == 0, the layout of POINT_jacobian
// (ag-build/cl/ec.cl:10-14) that ag_cuda_ec::multiple_multiexp copies into Vec<Projective>.
//
// Edge-case behaviour that is deliberate:
//   * an affine input (0,0) is the identity (ag-types/src/impls.rs:51-57) and is skipped; the
//     reference kernel silently computes garbage for it, its CPU path returns an error;
//   * P + (-P) yields infinity, P + P takes the doubling formula.
//
// Lazy-reduction contract (only FieldU29 cares; for FieldSat every hook is a no-op).  Notation
// "v<k": value < k*p; "l<k": every limb < k*2^29.  Invariant of every stored XYZZ point:
//     X: v<8 l<=1     Y: v<4 l<=1     ZZ, ZZZ: v<8 l<=1     infinity <=> ZZ has all limbs zero.
// mul/sqr need 9*la*lb + 9 < 64 and return v < va*vb/169 + 1, l<=1.  sub<K,LM>(a,b) = a - b + K p
// needs l_b <= LM and (K p >> 232) - LM >= b.v[8]; it returns v < va + K, l < la + LM + 1.
// Each line below carries its bound; tests/test_host_arith.py runs these formulas on the host with
// MSM_CHECK_BOUNDS (aborts on any 64-bit column overflow or negative limb) on worst-case inputs.
#pragma once
#include "fp.cuh"
#include "fp29.cuh"
#include "fp2.cuh"

namespace msm {

template <class F> struct Xyzz {
  typename F::Elem x, y, zz, zzz;
};
// Affine point in registers: canonical coordinates (v<1 l<=1) as unpacked from the resident copy
template <class F> struct Affine {
  typename F::Elem x, y;
};
// API-side layouts (little-endian 32-bit words, Montgomery R = 2^(32 N))
template <class F> struct ApiAffine {
  uint32_t x[F::API_WORDS], y[F::API_WORDS];
};
template <class F> struct ApiJacobian {
  uint32_t x[F::API_WORDS], y[F::API_WORDS], z[F::API_WORDS];
};
// Resident base point: what msm_bases_upload leaves in HBM (FieldSat: the API words unchanged;
// FieldU29: canonical x*2^261 mod p, re-sliced into 29-bit limbs on load)
template <class F> struct PackedAffine {
  uint32_t x[F::PACKED_WORDS], y[F::PACKED_WORDS];
};

template <class F> MSM_HD bool aff_is_identity(const Affine<F>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < F::N; i++) o |= a.x.v[i] | a.y.v[i];
  return o == 0;
}
// (x, -y) when negate, branch-free.  y canonical in; out y: v<=2 l<2 (lazy) -- fine as a mul
// operand against an l<=1 value, normalised where it is stored.
template <class F> MSM_HD Affine<F> aff_cneg(const Affine<F>& a, bool negate) {
  typename F::Elem n = F::template neg<2, 1>(a.y);
  Affine<F> r;
  r.x = a.x;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.y.v[i] = negate ? n.v[i] : a.y.v[i];
  return r;
}

template <class F> MSM_HD Xyzz<F> xyzz_inf() {
  Xyzz<F> r;
  r.x = F::zero();
  r.y = F::zero();
  r.zz = F::zero();
  r.zzz = F::zero();
  return r;
}
template <class F> MSM_HD bool xyzz_is_inf(const Xyzz<F>& a) { return F::is_zero_limbs(a.zz); }

// 2 * (affine point), mdbl-2008-s.  a.x canonical, a.y v<=2 l<2.
template <class F> MSM_COLD Xyzz<F> xyzz_mdbl(const Affine<F>& a) {
  using E = typename F::Elem;
  Xyzz<F> r;
  const E y = F::norm(a.y);                           // v<=2 l<=1
  const E u = F::add(y, y);                           // v<=4 l<2
  const E v = F::sqr(u);                              // (2,2)  v<1.1
  const E w = F::mul(u, v);                           // (2,1)  v<1.03
  const E s = F::mul(a.x, v);                         // v<1.01
  const E xx = F::sqr(a.x);                           // v<1.01
  const E m = F::norm(F::add(F::add(xx, xx), xx));    // v<3.03 l<=1
  const E x3 = F::norm(F::template sub<3, 2>(F::sqr(m), F::add(s, s)));  // v<1.06+3
  const E d = F::template sub<5, 1>(s, x3);           // v<6.01 l<3
  r.x = x3;
  r.y = F::norm(F::mul_sub(m, d, w, y));                               // (1,3),(1,1)  v<1.11+2
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * (xyzz point), dbl-2008-s-1 with a = 0
template <class F> MSM_COLD Xyzz<F> xyzz_dbl(const Xyzz<F>& a) {
  using E = typename F::Elem;
  if (xyzz_is_inf<F>(a)) return a;
  Xyzz<F> r;
  const E u = F::add(a.y, a.y);                       // v<8 l<2
  const E v = F::sqr(u);                              // (2,2)  v<1.38
  const E w = F::mul(u, v);                           // (2,1)  v<1.07
  const E s = F::mul(a.x, v);                         // v<1.07
  const E xx = F::sqr(a.x);                           // v<1.38
  const E m = F::norm(F::add(F::add(xx, xx), xx));    // v<4.2 l<=1
  const E x3 = F::norm(F::template sub<3, 2>(F::sqr(m), F::add(s, s)));  // v<1.11+3
  const E d = F::template sub<5, 1>(s, x3);           // v<6.1 l<3
  r.x = x3;
  r.y = F::norm(F::mul_sub(m, d, w, a.y));                             // v<1.16+2
  r.zz = F::mul(v, a.zz);
  r.zzz = F::mul(w, a.zzz);
  return r;
}

// acc += b, madd-2008-s: 8M + 2S.  b is not the identity encoding (callers filter (0,0));
// b.x canonical, b.y canonical or its lazy negation (v<=2 l<2).
template <class F> MSM_HD void xyzz_madd(Xyzz<F>& acc, const Affine<F>& b) {
  using E = typename F::Elem;
  if (xyzz_is_inf<F>(acc)) {
    acc.x = b.x;
    acc.y = F::norm(b.y);
    acc.zz = F::one();
    acc.zzz = F::one();
    return;
  }
  const E u2 = F::mul(b.x, acc.zz);                   // (1,1)  v<1.05
  const E s2 = F::mul(b.y, acc.zzz);                  // (2,1)  v<1.1
  const E pp_ = F::norm(F::template sub<9, 1>(u2, acc.x));   // P: v in (1, 10.05) p  l<=1
  const E r = F::norm(F::template sub<5, 1>(s2, acc.y));     // R: v in (1, 6.1) p   l<=1
  if (F::template is_multiple_of_p<1, 11>(pp_)) {     // U2 == X1: same x
    if (F::template is_multiple_of_p<1, 7>(r)) {
      acc = xyzz_mdbl<F>(b);                          // same point: double
    } else {
      acc = xyzz_inf<F>();                            // opposite points
    }
    return;
  }
  const E pp = F::sqr(pp_);                           // v<1.6
  const E ppp = F::mul(pp_, pp);                      // v<1.1
  const E q = F::mul(acc.x, pp);                      // v<1.08
  const E s = F::add(F::add(ppp, q), q);              // v<3.26 l<3
  const E x3 = F::norm(F::template sub<5, 3>(F::sqr(r), s));  // v<1.22+5
  const E d = F::template sub<7, 1>(q, x3);           // v<8.08 l<3
  const E y3 = F::norm(F::mul_sub(r, d, acc.y, ppp));  // r*d - Y1*PPP, one reduction; (1,3),(1,1)  v<1.29+2
  acc.x = x3;
  acc.y = y3;
  acc.zz = F::mul(acc.zz, pp);                        // v<1.08
  acc.zzz = F::mul(acc.zzz, ppp);                     // v<1.06
}

// a + b, add-2008-s: 12M + 2S.  xyzz_add_inl is the body, inlined where the addition IS the hot loop (the running sums
// of k_bucket_reduce: two additions per bucket; as an out-of-line call every operand goes through local memory);
// xyzz_add is the one out-of-line copy per curve everything else calls.
template <class F> MSM_HD Xyzz<F> xyzz_add_inl(const Xyzz<F>& a, const Xyzz<F>& b) {
  using E = typename F::Elem;
  if (xyzz_is_inf<F>(a)) return b;
  if (xyzz_is_inf<F>(b)) return a;
  const E u1 = F::mul(a.x, b.zz);                     // v<1.38
  const E u2 = F::mul(b.x, a.zz);                     // v<1.38
  const E s1 = F::mul(a.y, b.zzz);                    // v<1.19
  const E s2 = F::mul(b.y, a.zzz);                    // v<1.19
  const E pp_ = F::norm(F::template sub<2, 1>(u2, u1));      // v in (0.6, 3.4) p
  const E r = F::norm(F::template sub<2, 1>(s2, s1));        // v in (0.8, 3.2) p
  if (F::template is_multiple_of_p<0, 4>(pp_)) {
    if (F::template is_multiple_of_p<0, 4>(r)) return xyzz_dbl<F>(a);
    return xyzz_inf<F>();
  }
  const E pp = F::sqr(pp_);                           // v<1.07
  const E ppp = F::mul(pp_, pp);                      // v<1.03
  const E q = F::mul(u1, pp);                         // v<1.01
  const E s = F::add(F::add(ppp, q), q);              // v<3.1 l<3
  Xyzz<F> o;
  o.x = F::norm(F::template sub<5, 3>(F::sqr(r), s)); // v<1.07+5
  const E d = F::template sub<7, 1>(q, o.x);          // v<8.1 l<3
  o.y = F::norm(F::mul_sub(r, d, s1, ppp));                           // v<1.16+2
  o.zz = F::mul(F::mul(a.zz, b.zz), pp);
  o.zzz = F::mul(F::mul(a.zzz, b.zzz), ppp);
  return o;
}
template <class F> MSM_COLD Xyzz<F> xyzz_add(const Xyzz<F>& a, const Xyzz<F>& b) { return xyzz_add_inl<F>(a, b); }

// XYZZ -> Jacobian with Z = ZZZ:  X' = X*ZZ^2, Y' = Y*ZZZ^2, Z' = ZZZ
// (x = X'/Z'^2 = X ZZ^2 / ZZ^3 = X/ZZ;  y = Y'/Z'^3 = Y ZZZ^2 / ZZZ^3 = Y/ZZZ), written in the
// API layout.  Infinity -> POINT_ZERO = (0, 1, 0) (ag-build/cl/ec.cl:3).
template <class F> MSM_COLD void xyzz_to_api_jacobian(const Xyzz<F>& a, ApiJacobian<F>* out) {
  using E = typename F::Elem;
  if (xyzz_is_inf<F>(a)) {
    F::to_api(F::zero(), out->x);
    F::to_api(F::one(), out->y);
    F::to_api(F::zero(), out->z);
    return;
  }
  const E zz2 = F::sqr(a.zz);
  const E zzz2 = F::sqr(a.zzz);
  F::to_api(F::mul(a.x, zz2), out->x);
  F::to_api(F::mul(a.y, zzz2), out->y);
  F::to_api(a.zzz, out->z);
}

template <class F> MSM_COLD Xyzz<F> xyzz_from_api_jacobian(const ApiJacobian<F>* j) {
  using E = typename F::Elem;
  const E z = F::from_api(j->z);
  bool z_zero = true;
  for (int i = 0; i < F::API_WORDS; i++) z_zero = z_zero && (j->z[i] == 0);
  if (z_zero) return xyzz_inf<F>();
  Xyzz<F> r;
  r.x = F::from_api(j->x);
  r.y = F::from_api(j->y);
  r.zz = F::sqr(z);
  r.zzz = F::mul(r.zz, z);
  return r;
}

template <class F> MSM_COLD Affine<F> affine_from_api(const ApiAffine<F>* a) {
  Affine<F> r;
  bool ident = true;
  for (int i = 0; i < F::API_WORDS; i++) ident = ident && (a->x[i] == 0) && (a->y[i] == 0);
  if (ident) {
    r.x = F::zero();
    r.y = F::zero();
    return r;
  }
  r.x = F::norm(F::from_api(a->x));
  r.y = F::norm(F::from_api(a->y));
  return r;
}

// XYZZ -> affine, API layout: Montgomery (mont_out) or canonical integers; infinity -> (0,0).
// Returns true when the point is infinity.
template <class F> MSM_COLD bool xyzz_to_api_affine(const Xyzz<F>& a, bool mont_out, ApiAffine<F>* out) {
  using E = typename F::Elem;
  if (xyzz_is_inf<F>(a)) {
    for (int i = 0; i < F::API_WORDS; i++) out->x[i] = out->y[i] = 0;
    return true;
  }
  // 1/ZZ and 1/ZZZ from one inversion: i = 1/(ZZ*ZZZ); 1/ZZ = i*ZZZ; 1/ZZZ = i*ZZ
  const E i = F::inv(F::mul(a.zz, a.zzz));
  const E x = F::mul(a.x, F::mul(i, a.zzz));
  const E y = F::mul(a.y, F::mul(i, a.zz));
  if (mont_out) {
    F::to_api(x, out->x);
    F::to_api(y, out->y);
  } else {
    F::to_canonical_words(x, out->x);
    F::to_canonical_words(y, out->y);
  }
  return false;
}

// k * b for a small unsigned k (bucket weights), double-and-add MSB first
template <class F> MSM_COLD Xyzz<F> xyzz_mul_small(const Xyzz<F>& b, uint32_t k) {
  Xyzz<F> acc = xyzz_inf<F>();
  for (int i = 31; i >= 0; i--) {
    acc = xyzz_dbl<F>(acc);
    if ((k >> i) & 1) acc = xyzz_add<F>(acc, b);
  }
  return acc;
}

}  // namespace msm
