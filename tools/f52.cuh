// f52.cuh -- Montgomery product on the FP64 pipe: N52 limbs of 52 bits, R' = 2^(52 N52).
//
// B200 issues DFMA at the same rate as IMAD (64 / clk / SM, tools/pipe_model.cu) on a separate
// pipe, while the carry-chained IMAD.WIDE.U32.X that a 32-bit-limb product is made of runs at half
// rate.  This file is the product formulated for that pipe (after Emmart, Zheng, Weems, "Faster
// modular exponentiation using double precision floating point arithmetic on the GPU", ARITH 2018):
// for integer-valued doubles a, b < 2^52
//     h = fma_rz(a, b, 2^104)              = 2^104 + floor(a b / 2^52) 2^52      (exact, ulp 2^52)
//     l = fma_rz(a, b, (2^104 + 2^52) - h) = 2^52 + (a b mod 2^52)               (exact)
// so the raw bit patterns of h and l are a constant plus the high / low half of the product; they
// are summed per column as 64-bit integers and the constants are taken out once per column.  The
// constants have zero low 52 bits, so the Montgomery quotient digit can be read off a column at
// any time.  3 FP64 operations and two 64-bit integer additions per 52 x 52 partial product;
// 2 N52^2 + N52 partial products per field product (N52 = 5 for BN254: 55).
//
// Replaces nothing in the reference (its field layer is 32/64-bit integer only,
// ag-build/cl/field.cl:85-299); it is the second arithmetic pipe of the accumulate kernel.
#pragma once
#include "../0g-ec-gpu_b200/csrc/ptx.cuh"

namespace msm {

constexpr uint64_t M52 = (1ull << 52) - 1;

struct Bn254Fq52 {
  static constexpr int N = 5;
  static constexpr int API_WORDS = 8;  // 32-bit words at the API boundary, Montgomery R = 2^256
  static MSM_HD constexpr uint64_t P(int i) {
    constexpr uint64_t t[N] = {0x08c16d87cfd47ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull,
                               0x030644e72e131ull};
    return t[i];
  }
  static constexpr uint64_t PINV = 0x20782e4866389ull;  // -p^-1 mod 2^52
  static MSM_HD constexpr uint64_t ONE(int i) {          // 2^260 mod p
    constexpr uint64_t t[N] = {0x20880f6fce4b4ull, 0x49baa989a8455ull, 0x18f014a498908ull, 0x724f85a9201d8ull,
                               0x01f16424e1bb7ull};
    return t[i];
  }
  static MSM_HD constexpr uint64_t FROM_API(int i) {     // 2^264 mod p:  mont'(x 2^256, .) = x 2^260
    constexpr uint64_t t[N] = {0xb0f2afaec667aull, 0xed9626b0fffbdull, 0x9e2a0fcad825aull, 0xe357276f48b70ull,
                               0x00d791464ef86ull};
    return t[i];
  }
  static MSM_HD constexpr uint64_t TO_API(int i) {       // 2^256 mod p:  mont'(x 2^260, .) = x 2^256
    constexpr uint64_t t[N] = {0xd438dc58f0d9dull, 0x28f5c70b3dd35ull, 0x879462c0a78ebull, 0xdf2f666ea36f7ull,
                               0x00e0a77c19a07ull};
    return t[i];
  }
};

struct Bls381Fq52 {
  static constexpr int N = 8;
  static constexpr int API_WORDS = 12;  // Montgomery R = 2^384
  static MSM_HD constexpr uint64_t P(int i) {
    constexpr uint64_t t[N] = {0xeffffffffaaabull, 0xfeb153ffffb9full, 0x6b0f6241eabffull, 0x12bf6730d2a0full,
                               0x764774b84f385ull, 0x1ba7b6434bacdull, 0x1ea397fe69a4bull, 0x000000001a011ull};
    return t[i];
  }
  static constexpr uint64_t PINV = 0x3fffcfffcfffdull;
  static MSM_HD constexpr uint64_t ONE(int i) {  // 2^416 mod p
    constexpr uint64_t t[N] = {0x6480ea8e9b9afull, 0x65766c8fe444full, 0x8b540fea96f7dull, 0x3b2ee82efd422ull,
                               0xa6723e5f0ade5ull, 0xff6eb6fdd4230ull, 0xe06ef23c24a25ull, 0x0000000014c8eull};
    return t[i];
  }
  static MSM_HD constexpr uint64_t FROM_API(int i) {  // 2^448 mod p
    constexpr uint64_t t[N] = {0x7fde37dba9366ull, 0x4e27525bc342bull, 0x1f5b1e9778489ull, 0xb872b2b91b9dcull,
                               0xb206f497dfcafull, 0x4137cc89a9b0bull, 0xd9d20d7e39959ull, 0x000000000411cull};
    return t[i];
  }
  static MSM_HD constexpr uint64_t TO_API(int i) {  // 2^384 mod p
    constexpr uint64_t t[N] = {0x900000002fffdull, 0x0bc40c0002760ull, 0x3c758baebf400ull, 0x57455f4898575ull,
                               0xd77ce58537052ull, 0x071a97a256ec6ull, 0xec3fa80e4935cull, 0x0000000015f65ull};
    return t[i];
  }
};

// integer < 2^52 -> the same value as a double (exact): one logic op on the high word, one DADD
MSM_HD double u52_to_double(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)(x | 0x4330000000000000ull)) - 4503599627370496.0;
#else
  return (double)x;
#endif
}

// number of (i, j) in [0, N)^2 with i + j == k
MSM_HD constexpr int f52_pairs(int N, int k) { return k < 0 || k > 2 * N - 2 ? 0 : (k < N ? k + 1 : 2 * N - 1 - k); }
// what the bit-pattern constants of all h / l terms of one product add up to in column k (mod 2^64):
// 2 n(k) low terms and 2 n(k-1) high terms (a*b and q*p)
MSM_HD constexpr uint64_t f52_column_bias(int N, int k) {
  return 2ull * (uint64_t)f52_pairs(N, k) * 0x4330000000000000ull + 2ull * (uint64_t)f52_pairs(N, k - 1) * 0x4670000000000000ull;
}

// col_lo += (a b mod 2^52), col_hi += floor(a b / 2^52)   (device: plus the bit-pattern constants)
MSM_HD void f52_term(double a, double b, uint64_t& col_lo, uint64_t& col_hi) {
#if defined(__CUDA_ARCH__)
  const double h = __fma_rz(a, b, 20282409603651670423947251286016.0);                      // 2^104
  const double l = __fma_rz(a, b, 20282409603651674927546878656512.0 - h);                  // 2^104 + 2^52
  col_hi += (uint64_t)__double_as_longlong(h);
  col_lo += (uint64_t)__double_as_longlong(l);
#else
  const unsigned __int128 pr = (unsigned __int128)(uint64_t)a * (uint64_t)b;
  col_hi += (uint64_t)(pr >> 52);
  col_lo += (uint64_t)pr & M52;
#endif
}
// (t * PINV) mod 2^52 as a double; t < 2^52
template <class Q> MSM_HD double f52_quotient(uint64_t t) {
#if defined(__CUDA_ARCH__)
  const double td = u52_to_double(t);
  const double h = __fma_rz(td, (double)Q::PINV, 20282409603651670423947251286016.0);
  const double l = __fma_rz(td, (double)Q::PINV, 20282409603651674927546878656512.0 - h);
  return l - 4503599627370496.0;
#else
  return (double)((uint64_t)((unsigned __int128)t * Q::PINV) & M52);
#endif
}

// r = a b / 2^(52 N) mod p, not fully reduced: r < a b / R' + p.  a, b: integer-valued doubles,
// every limb < 2^52.  r: limbs < 2^52.
template <class Q> MSM_HD void f52_mul_core(uint64_t* r, const double* a, const double* b) {
  constexpr int N = Q::N;
  uint64_t c[2 * N];
#pragma unroll
  for (int k = 0; k < 2 * N; k++) {
#if defined(__CUDA_ARCH__)
    c[k] = 0ull - f52_column_bias(N, k);
#else
    c[k] = 0;
#endif
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int j = 0; j < N; j++) f52_term(a[j], b[i], c[i + j], c[i + j + 1]);
    const double q = f52_quotient<Q>(c[i] & M52);
#pragma unroll
    for (int j = 0; j < N; j++) f52_term(q, (double)Q::P(j), c[i + j], c[i + j + 1]);
    c[i + 1] += c[i] >> 52;
  }
#pragma unroll
  for (int k = N; k < 2 * N - 1; k++) {
    c[k + 1] += c[k] >> 52;
    r[k - N] = c[k] & M52;
  }
  r[N - 1] = c[2 * N - 1];
}

}  // namespace msm
