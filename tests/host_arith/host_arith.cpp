// Host build of csrc/fp.cuh + fp29.cuh + ec.cuh (emulated carry flag) so the field and curve
// formulas the kernels use can be checked against the oracle on a box without a GPU, with the
// lazy-reduction bound assertions switched on (-DMSM_CHECK_BOUNDS).  Test vehicle only.
#include <cstddef>
#include <cstring>
#include "../../0g-ec-gpu_b200/csrc/ec.cuh"
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
