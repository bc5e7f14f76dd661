// Host build of csrc/fp.cuh + csrc/ec.cuh (emulated carry flag) so the field and curve formulas
// the kernels use can be checked against the oracle on a box without a GPU.  Test vehicle only.
#include <cstddef>
#include <cstring>
#include "../../0g-ec-gpu_b200/csrc/ec.cuh"
using namespace msm;

template <class P> static int fq_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) {
  constexpr int N = P::N;
  for (size_t i = 0; i < n; i++) {
    Fp<P> x, y, r;
    memcpy(x.v, a + i * N, 4 * N);
    if (b) memcpy(y.v, b + i * N, 4 * N);
    switch (op) {
      case 0: r = fp_add<P>(x, y); break;
      case 1: r = fp_sub<P>(x, y); break;
      case 2: r = fp_mul<P>(x, y); break;
      case 3: r = fp_sqr<P>(x); break;
      case 4: r = fp_dbl<P>(x); break;
      case 5: r = fp_to_mont<P>(x); break;
      case 6: r = fp_from_mont<P>(x); break;
      case 7: r = fp_inv<P>(x); break;
      case 8: r = fp_neg<P>(x); break;
      default: return -1;
    }
    memcpy(o + i * N, r.v, 4 * N);
  }
  return 0;
}
// op: 0 = add(Jac,Jac) via xyzz_add, 1 = madd(Jac, Aff) via xyzz_madd, 2 = dbl(Jac) via xyzz_dbl,
//     3 = mdbl(Aff b) , 4 = to_affine(Jac a) -> writes {x,y,0}, 5 = mul_small(Jac a, k = b[0])
template <class P> static int ec_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) {
  constexpr int N = P::N;
  for (size_t i = 0; i < n; i++) {
    Jacobian<P> ja, jb;
    Affine<P> ab;
    memcpy(&ja, a + i * 3 * N, 12 * N);
    Xyzz<P> xa = xyzz_from_jacobian<P>(ja), r;
    switch (op) {
      case 0: memcpy(&jb, b + i * 3 * N, 12 * N); r = xyzz_add<P>(xa, xyzz_from_jacobian<P>(jb)); break;
      case 1: memcpy(&ab, b + i * 2 * N, 8 * N); r = xa; if (!aff_is_identity<P>(ab)) xyzz_madd<P>(r, ab); break;
      case 2: r = xyzz_dbl<P>(xa); break;
      case 3: memcpy(&ab, b + i * 2 * N, 8 * N); r = xyzz_mdbl<P>(ab); break;
      case 4: { Affine<P> af = xyzz_to_affine<P>(xa); memset(o + i * 3 * N, 0, 12 * N); memcpy(o + i * 3 * N, &af, 8 * N); continue; }
      case 5: r = xyzz_mul_small<P>(xa, b[i]); break;
      default: return -1;
    }
    Jacobian<P> jo = xyzz_to_jacobian<P>(r);
    memcpy(o + i * 3 * N, &jo, 12 * N);
  }
  return 0;
}
extern "C" int host_fq_op(int curve, int op, const void* a, const void* b, void* o, size_t n) {
  if (curve == 0) return fq_op<Bn254Fq>(op, (const uint32_t*)a, (const uint32_t*)b, (uint32_t*)o, n);
  if (curve == 1) return fq_op<Bls381Fq>(op, (const uint32_t*)a, (const uint32_t*)b, (uint32_t*)o, n);
  return -100;
}
extern "C" int host_ec_op(int curve, int op, const void* a, const void* b, void* o, size_t n) {
  if (curve == 0) return ec_op<Bn254Fq>(op, (const uint32_t*)a, (const uint32_t*)b, (uint32_t*)o, n);
  if (curve == 1) return ec_op<Bls381Fq>(op, (const uint32_t*)a, (const uint32_t*)b, (uint32_t*)o, n);
  return -100;
}
