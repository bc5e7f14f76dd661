// Host build of csrc/fp.cuh + fp29.cuh + ec.cuh (emulated carry flag) so the field and curve
// formulas the kernels use can be checked against the oracle on a box without a GPU, with the
// lazy-reduction bound assertions switched on (-DMSM_CHECK_BOUNDS).  Test vehicle only.
#include <cstddef>
#include <cstring>
#include <vector>
#include "../../0g-ec-gpu_b200/csrc/ec.cuh"
#include "../../0g-ec-gpu_b200/csrc/bucket_affine.cuh"
using namespace msm;

template <class F> static int fq_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* r2, uint32_t* o, size_t n) {
  using E = typename F::Elem;
  constexpr int W = F::API_WORDS;
  for (size_t i = 0; i < n; i++) {
    E x = F::norm(F::from_api(a + i * W));
    E y = b ? F::norm(F::from_api(b + i * W)) : x;
    E r;
    switch (op) {
      case 0: r = F::add(x, y); break;
      case 1: r = F::template sub<2, 1>(x, y); break;
      case 2: r = F::mul(x, y); break;
      case 3: r = F::sqr(x); break;
      case 4: r = F::add(x, x); break;
      case 5: r = F::mul(x, F::from_api(r2)); break;
      case 6: { uint32_t unit[W] = {1}; r = F::mul(x, F::from_api(unit)); break; }
      case 7: r = F::inv(x); break;
      case 8: r = F::template neg<2, 1>(x); break;
      default: return -1;
    }
    F::to_api(r, o + i * W);
  }
  return 0;
}
// op: 0 add(Jac,Jac)  1 madd(Jac,Aff)  2 dbl(Jac)  3 mdbl(Aff b)  4 to_affine(Jac a, Montgomery) -> {x,y,0}
//     5 mul_small(Jac a, k = b[i])  6 madd with negated b  7 api_to_packed + unpack round trip of b (Aff) -> {x,y,0}
template <class F> static int ec_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) {
  constexpr int W = F::API_WORDS;
  for (size_t i = 0; i < n; i++) {
    const ApiJacobian<F>* ja = reinterpret_cast<const ApiJacobian<F>*>(a) + i;
    ApiJacobian<F>* jo = reinterpret_cast<ApiJacobian<F>*>(o) + i;
    Xyzz<F> xa = xyzz_from_api_jacobian<F>(ja), r;
    switch (op) {
      case 0: r = xyzz_add<F>(xa, xyzz_from_api_jacobian<F>(reinterpret_cast<const ApiJacobian<F>*>(b) + i)); break;
      case 1:
      case 6: {
        Affine<F> q = affine_from_api<F>(reinterpret_cast<const ApiAffine<F>*>(b) + i);
        r = xa;
        if (!aff_is_identity<F>(q)) { q = aff_cneg<F>(q, op == 6); xyzz_madd<F>(r, q); }
        break;
      }
      case 2: r = xyzz_dbl<F>(xa); break;
      case 3: { Affine<F> q = affine_from_api<F>(reinterpret_cast<const ApiAffine<F>*>(b) + i); r = xyzz_mdbl<F>(q); break; }
      case 4: {
        memset(jo, 0, sizeof(*jo));
        xyzz_to_api_affine<F>(xa, true, reinterpret_cast<ApiAffine<F>*>(jo));
        continue;
      }
      case 5: r = xyzz_mul_small<F>(xa, b[i]); break;
      case 7: {
        const ApiAffine<F>* q = reinterpret_cast<const ApiAffine<F>*>(b) + i;
        PackedAffine<F> pk;
        F::api_to_packed(q->x, pk.x);
        F::api_to_packed(q->y, pk.y);
        memset(jo, 0, sizeof(*jo));
        F::to_api(F::unpack(pk.x), jo->x);
        F::to_api(F::unpack(pk.y), jo->y);
        continue;
      }
      default: return -1;
    }
    xyzz_to_api_jacobian<F>(r, jo);
  }
  (void)W;
  return 0;
}
// a chain of `steps` mixed additions acc += (+-)pts[(i*stride + k) % m] starting from infinity, per lane i:
// exercises the accumulator invariants of xyzz_madd over long runs
template <class F> static int madd_chain(const uint32_t* pts, size_t m, size_t steps, size_t lanes, uint32_t* o) {
  for (size_t i = 0; i < lanes; i++) {
    Xyzz<F> acc = xyzz_inf<F>();
    for (size_t k = 0; k < steps; k++) {
      size_t idx = (i * 7 + k * (i + 1)) % m;
      Affine<F> q = affine_from_api<F>(reinterpret_cast<const ApiAffine<F>*>(pts) + idx);
      if (aff_is_identity<F>(q)) continue;
      q = aff_cneg<F>(q, ((k ^ i) & 1) != 0);
      xyzz_madd<F>(acc, q);
    }
    xyzz_to_api_jacobian<F>(acc, reinterpret_cast<ApiJacobian<F>*>(o) + i);
  }
  return 0;
}

#define DISPATCH(CALL)                                                     \
  if (curve == 0 && impl == 0) return CALL(FieldSat<Bn254Fq>);              \
  if (curve == 0 && impl == 1) return CALL(FieldU29<Bn254U29>);             \
  if (curve == 1 && impl == 0) return CALL(FieldSat<Bls381Fq>);             \
  if (curve == 0 && impl == 2) { using L0 = FieldSatLazy<Bn254Fq>; return CALL(L0); }   \
  if (curve == 1 && impl == 2) { using L1 = FieldSatLazy<Bls381Fq>; return CALL(L1); }  \
  if (curve == 2 && impl == 2) { using E0 = FieldExt2Lazy<Bn254Fq>; return CALL(E0); }   \
  if (curve == 3 && impl == 2) { using E1 = FieldExt2Lazy<Bls381Fq>; return CALL(E1); }  \
  return -100;

extern "C" int host_fq_op(int curve, int impl, int op, const void* a, const void* b, const void* r2, void* o, size_t n) {
#define CALL(F) fq_op<F>(op, (const uint32_t*)a, (const uint32_t*)b, (const uint32_t*)r2, (uint32_t*)o, n)
  DISPATCH(CALL)
#undef CALL
}
extern "C" int host_ec_op(int curve, int impl, int op, const void* a, const void* b, void* o, size_t n) {
#define CALL(F) ec_op<F>(op, (const uint32_t*)a, (const uint32_t*)b, (uint32_t*)o, n)
  DISPATCH(CALL)
#undef CALL
}
extern "C" int host_madd_chain(int curve, int impl, const void* pts, size_t m, size_t steps, size_t lanes, void* o) {
#define CALL(F) madd_chain<F>((const uint32_t*)pts, m, steps, lanes, (uint32_t*)o)
  DISPATCH(CALL)
#undef CALL
}

// Affine halving rounds (csrc/bucket_affine.cuh), the per-thread device functions run thread by thread:
// `rounds` rounds over the sorted entry list, T emulated threads, batches of at most m_max items.
// (The kernel shares one inversion per block through a shared-memory product tree; here every thread
// inverts its own batch product -- the same value.)
// bases: API-layout affine points; entries: index | sign << 31, grouped by bucket (off0[NB + 1]).
// Returns the number of points left; out_pts (API affine layout) / out_off ([NB + 1]) describe them.
template <class F>
static long ba_rounds(const uint32_t* bases_api, size_t n_bases, const uint32_t* entries, const uint32_t* off0, uint32_t NB,
                      uint32_t rounds, uint32_t T, uint32_t m_max, uint32_t* out_pts, uint32_t* out_off) {
  constexpr int N = F::N;
  std::vector<PackedAffine<F>> packed(n_bases);
  for (size_t i = 0; i < n_bases; i++) {
    const ApiAffine<F>* a = reinterpret_cast<const ApiAffine<F>*>(bases_api) + i;
    F::api_to_packed(a->x, packed[i].x);
    F::api_to_packed(a->y, packed[i].y);
  }
  std::vector<uint32_t> off_in(off0, off0 + NB + 1), off_out(NB + 1);
  std::vector<uint4> cx, cy, nx, ny;  // planes (16-byte aligned)
  std::vector<uint4> sc_prefix((size_t)T * m_max * N / 4);
  std::vector<uint32_t> sc_idx((size_t)T * m_max);
  BaPoints<F> in{packed.data(), entries, nullptr, nullptr};
  for (uint32_t r = 0; r < rounds; r++) {
    off_out[0] = 0;
    for (uint32_t g = 0; g < NB; g++) off_out[g + 1] = off_out[g] + ((off_in[g + 1] - off_in[g] + 1) >> 1);
    const size_t cap = (size_t)off_out[NB] + 1;
    nx.assign(cap * N / 4, uint4{0, 0, 0, 0});
    ny.assign(cap * N / 4, uint4{0, 0, 0, 0});
    BaPoints<F> out{nullptr, nullptr, reinterpret_cast<uint32_t*>(nx.data()), reinterpret_cast<uint32_t*>(ny.data())};
    const BaGeom gm = ba_geom(T, off_out[NB]);
    std::vector<BaWalk> wk(T);
    for (auto& w : wk) w.ready = false;
    uint32_t* pre = reinterpret_cast<uint32_t*>(sc_prefix.data());
    for (uint32_t i0 = 0; i0 < gm.per; i0 += m_max) {
      std::vector<typename F::Elem> prod(T);
      for (uint32_t t = 0; t < T; t++)
        prod[t] = r == 0 ? ba_forward<F, true>(t, gm, i0, m_max, in, off_in.data(), off_out.data(), NB, wk[t], pre, sc_idx.data())
                         : ba_forward<F, false>(t, gm, i0, m_max, in, off_in.data(), off_out.data(), NB, wk[t], pre, sc_idx.data());
      for (uint32_t t = 0; t < T; t++) {
        const typename F::Elem inv = F::inv(prod[t]);
        if (r == 0) ba_backward<F, true>(t, gm, i0, m_max, in, out, inv, pre, sc_idx.data());
        else ba_backward<F, false>(t, gm, i0, m_max, in, out, inv, pre, sc_idx.data());
      }
    }
    cx.swap(nx);
    cy.swap(ny);
    in = BaPoints<F>{nullptr, nullptr, reinterpret_cast<uint32_t*>(cx.data()), reinterpret_cast<uint32_t*>(cy.data())};
    off_in = off_out;
  }
  const uint32_t left = off_in[NB];
  for (uint32_t i = 0; i < left; i++) {
    ApiAffine<F>* o = reinterpret_cast<ApiAffine<F>*>(out_pts) + i;
    F::to_api(F::unpack(in.x + (size_t)i * N), o->x);
    F::to_api(F::unpack(in.y + (size_t)i * N), o->y);
  }
  for (uint32_t g = 0; g <= NB; g++) out_off[g] = off_in[g];
  return (long)left;
}
extern "C" long host_ba_rounds(int curve, const void* bases, size_t n_bases, const void* entries, const void* off0,
                               uint32_t NB, uint32_t rounds, uint32_t T, uint32_t m_max, void* out_pts, void* out_off) {
  if (curve == 0)
    return ba_rounds<FieldSatLazy<Bn254Fq>>((const uint32_t*)bases, n_bases, (const uint32_t*)entries, (const uint32_t*)off0,
                                            NB, rounds, T, m_max, (uint32_t*)out_pts, (uint32_t*)out_off);
  if (curve == 1)
    return ba_rounds<FieldSatLazy<Bls381Fq>>((const uint32_t*)bases, n_bases, (const uint32_t*)entries, (const uint32_t*)off0,
                                             NB, rounds, T, m_max, (uint32_t*)out_pts, (uint32_t*)out_off);
  return -100;
}
