// ec.cuh -- short-Weierstrass a = 0 group law for the MSM buckets.
//
// Replaces POINT_add_mixed / POINT_add / POINT_double of ag-build/cl/ec.cl:17-120.  Buckets are
// kept in extended Jacobian ("XYZZ") coordinates, x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2 (EFD
// "shortw/xyzz": madd-2008-s = 8M+2S, add-2008-s = 12M+2S, dbl-2008-s-1, mdbl-2008-s) instead of
// the reference's Jacobian madd-2007-bl (7M+4S, executed as 11 multiplies).  Results leave the
// engine as Jacobian {x,y,z} Montgomery with infinity <=> z == 0, the layout of POINT_jacobian
// (ag-build/cl/ec.cl:10-14) that ag_cuda_ec::multiple_multiexp copies into Vec<Projective>.
//
// Differences in edge-case behaviour that are deliberate:
//   * an affine input (0,0) is the identity (ag-types/src/impls.rs:51-57) and is skipped; the
//     reference kernel silently computes garbage for it, its CPU path returns an error;
//   * P + (-P) yields infinity (the reference's madd-2007-bl would produce z = 0 as well).
#pragma once
#include "fp.cuh"

namespace msm {

template <class P> struct Affine {
  Fp<P> x, y;
};
template <class P> struct Xyzz {
  Fp<P> x, y, zz, zzz;
};
template <class P> struct Jacobian {
  Fp<P> x, y, z;
};

template <class P> MSM_HD bool aff_is_identity(const Affine<P>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) o |= a.x.v[i] | a.y.v[i];
  return o == 0;
}
template <class P> MSM_HD Affine<P> aff_neg(const Affine<P>& a) {
  Affine<P> r;
  r.x = a.x;
  r.y = fp_neg<P>(a.y);
  return r;
}
// conditional negation without a branch on the data path
template <class P> MSM_HD Affine<P> aff_cneg(const Affine<P>& a, bool negate) {
  Affine<P> n = aff_neg<P>(a);
  Affine<P> r;
  r.x = a.x;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.y.v[i] = negate ? n.y.v[i] : a.y.v[i];
  return r;
}

template <class P> MSM_HD Xyzz<P> xyzz_inf() {
  Xyzz<P> r;
  r.x = fp_zero<P>();
  r.y = fp_zero<P>();
  r.zz = fp_zero<P>();
  r.zzz = fp_zero<P>();
  return r;
}
template <class P> MSM_HD bool xyzz_is_inf(const Xyzz<P>& a) { return fp_is_zero<P>(a.zz); }

template <class P> MSM_HD Xyzz<P> xyzz_from_affine(const Affine<P>& a) {
  Xyzz<P> r;
  if (aff_is_identity<P>(a)) return xyzz_inf<P>();
  r.x = a.x;
  r.y = a.y;
  r.zz = fp_one<P>();
  r.zzz = fp_one<P>();
  return r;
}

// 2 * (affine point), mdbl-2008-s
template <class P> MSM_COLD Xyzz<P> xyzz_mdbl(const Affine<P>& a) {
  Xyzz<P> r;
  Fp<P> u = fp_dbl<P>(a.y);
  Fp<P> v = fp_sqr<P>(u);
  Fp<P> w = fp_mul<P>(u, v);
  Fp<P> s = fp_mul<P>(a.x, v);
  Fp<P> xx = fp_sqr<P>(a.x);
  Fp<P> m = fp_add<P>(fp_dbl<P>(xx), xx);
  r.x = fp_sub<P>(fp_sqr<P>(m), fp_dbl<P>(s));
  r.y = fp_sub<P>(fp_mul<P>(m, fp_sub<P>(s, r.x)), fp_mul<P>(w, a.y));
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * (xyzz point), dbl-2008-s-1 with a = 0
template <class P> MSM_COLD Xyzz<P> xyzz_dbl(const Xyzz<P>& a) {
  if (xyzz_is_inf<P>(a)) return a;
  Xyzz<P> r;
  Fp<P> u = fp_dbl<P>(a.y);
  Fp<P> v = fp_sqr<P>(u);
  Fp<P> w = fp_mul<P>(u, v);
  Fp<P> s = fp_mul<P>(a.x, v);
  Fp<P> xx = fp_sqr<P>(a.x);
  Fp<P> m = fp_add<P>(fp_dbl<P>(xx), xx);
  r.x = fp_sub<P>(fp_sqr<P>(m), fp_dbl<P>(s));
  r.y = fp_sub<P>(fp_mul<P>(m, fp_sub<P>(s, r.x)), fp_mul<P>(w, a.y));
  r.zz = fp_mul<P>(v, a.zz);
  r.zzz = fp_mul<P>(w, a.zzz);
  return r;
}

// acc += b (affine, not the identity encoding -- callers filter (0,0)), madd-2008-s: 8M + 2S
template <class P> MSM_HD void xyzz_madd(Xyzz<P>& acc, const Affine<P>& b) {
  if (xyzz_is_inf<P>(acc)) {
    acc.x = b.x;
    acc.y = b.y;
    acc.zz = fp_one<P>();
    acc.zzz = fp_one<P>();
    return;
  }
  Fp<P> u2 = fp_mul<P>(b.x, acc.zz);
  Fp<P> s2 = fp_mul<P>(b.y, acc.zzz);
  Fp<P> pp_ = fp_sub<P>(u2, acc.x);  // P
  Fp<P> r = fp_sub<P>(s2, acc.y);    // R
  if (fp_is_zero<P>(pp_)) {
    if (fp_is_zero<P>(r)) {
      acc = xyzz_mdbl<P>(b);
    } else {
      acc = xyzz_inf<P>();
    }
    return;
  }
  Fp<P> pp = fp_sqr<P>(pp_);
  Fp<P> ppp = fp_mul<P>(pp_, pp);
  Fp<P> q = fp_mul<P>(acc.x, pp);
  Fp<P> x3 = fp_sub<P>(fp_sub<P>(fp_sqr<P>(r), ppp), fp_dbl<P>(q));
  Fp<P> y3 = fp_sub<P>(fp_mul<P>(r, fp_sub<P>(q, x3)), fp_mul<P>(acc.y, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = fp_mul<P>(acc.zz, pp);
  acc.zzz = fp_mul<P>(acc.zzz, ppp);
}

// a + b, add-2008-s: 12M + 2S
template <class P> MSM_COLD Xyzz<P> xyzz_add(const Xyzz<P>& a, const Xyzz<P>& b) {
  if (xyzz_is_inf<P>(a)) return b;
  if (xyzz_is_inf<P>(b)) return a;
  Fp<P> u1 = fp_mul<P>(a.x, b.zz);
  Fp<P> u2 = fp_mul<P>(b.x, a.zz);
  Fp<P> s1 = fp_mul<P>(a.y, b.zzz);
  Fp<P> s2 = fp_mul<P>(b.y, a.zzz);
  Fp<P> pp_ = fp_sub<P>(u2, u1);
  Fp<P> r = fp_sub<P>(s2, s1);
  if (fp_is_zero<P>(pp_)) {
    if (fp_is_zero<P>(r)) return xyzz_dbl<P>(a);
    return xyzz_inf<P>();
  }
  Fp<P> pp = fp_sqr<P>(pp_);
  Fp<P> ppp = fp_mul<P>(pp_, pp);
  Fp<P> q = fp_mul<P>(u1, pp);
  Xyzz<P> o;
  o.x = fp_sub<P>(fp_sub<P>(fp_sqr<P>(r), ppp), fp_dbl<P>(q));
  o.y = fp_sub<P>(fp_mul<P>(r, fp_sub<P>(q, o.x)), fp_mul<P>(s1, ppp));
  o.zz = fp_mul<P>(fp_mul<P>(a.zz, b.zz), pp);
  o.zzz = fp_mul<P>(fp_mul<P>(a.zzz, b.zzz), ppp);
  return o;
}

// XYZZ -> Jacobian with Z = ZZZ:  X' = X*ZZ^2, Y' = Y*ZZZ^2 ( = Y*ZZ^3 ), Z' = ZZZ
// (x = X'/Z'^2 = X ZZ^2 / ZZ^3 = X/ZZ;  y = Y'/Z'^3 = Y ZZZ^2 / ZZZ^3 = Y/ZZZ).
// Infinity -> POINT_ZERO = (0, 1, 0) (ag-build/cl/ec.cl:3).
template <class P> MSM_COLD Jacobian<P> xyzz_to_jacobian(const Xyzz<P>& a) {
  Jacobian<P> j;
  if (xyzz_is_inf<P>(a)) {
    j.x = fp_zero<P>();
    j.y = fp_one<P>();
    j.z = fp_zero<P>();
    return j;
  }
  Fp<P> zz2 = fp_sqr<P>(a.zz);
  Fp<P> zzz2 = fp_sqr<P>(a.zzz);
  j.x = fp_mul<P>(a.x, zz2);
  j.y = fp_mul<P>(a.y, zzz2);
  j.z = a.zzz;
  return j;
}

template <class P> MSM_COLD Xyzz<P> xyzz_from_jacobian(const Jacobian<P>& j) {
  if (fp_is_zero<P>(j.z)) return xyzz_inf<P>();
  Xyzz<P> r;
  r.x = j.x;
  r.y = j.y;
  r.zz = fp_sqr<P>(j.z);
  r.zzz = fp_mul<P>(r.zz, j.z);
  return r;
}

// XYZZ -> affine (Montgomery); infinity -> (0,0)
template <class P> MSM_COLD Affine<P> xyzz_to_affine(const Xyzz<P>& a) {
  Affine<P> r;
  if (xyzz_is_inf<P>(a)) {
    r.x = fp_zero<P>();
    r.y = fp_zero<P>();
    return r;
  }
  // 1/ZZ and 1/ZZZ from one inversion: i = 1/(ZZ*ZZZ); 1/ZZ = i*ZZZ; 1/ZZZ = i*ZZ
  Fp<P> i = fp_inv<P>(fp_mul<P>(a.zz, a.zzz));
  r.x = fp_mul<P>(a.x, fp_mul<P>(i, a.zzz));
  r.y = fp_mul<P>(a.y, fp_mul<P>(i, a.zz));
  return r;
}

// k * b for a small unsigned k (bucket weights), double-and-add MSB first
template <class P> MSM_COLD Xyzz<P> xyzz_mul_small(const Xyzz<P>& b, uint32_t k) {
  Xyzz<P> acc = xyzz_inf<P>();
  for (int i = 31; i >= 0; i--) {
    acc = xyzz_dbl<P>(acc);
    if ((k >> i) & 1) acc = xyzz_add<P>(acc, b);
  }
  return acc;
}

}  // namespace msm
