// ecfft.cuh -- discrete Fourier transform of G1 points, on the same field
// and curve headers as the MSM.
//
// Replaces KERNEL POINT_radix_fft (ag-build/cl/ec-fft.cl:4-76) and its host drivers
// ag_cuda_ec::ec_fft::radix_ec_fft (ag-cuda-ec/src/ec_fft.rs:13-99) and
// SingleEcFftKernel::radix_ec_fft (ec-gpu-proxy/src/ec_fft.rs:53-160):
//     out[k] = sum_j omega^(j k) * in[j],   n = 2^log_n points, natural order in and out,
// the result Radix2EvaluationDomain::fft gives (ag-cuda-ec/src/ec_fft.rs:131-137).  The reference
// runs radix-2^deg passes with a full 256-bit double-and-add POINT_mul per twiddle and a SCALAR_pow
// per butterfly; here: one table of twiddles omega^j (canonical integers, one Fr product chain per
// entry), log_n decimation-in-time rounds of n/2 butterflies, one signed 4-bit-window scalar
// multiplication in XYZZ coordinates per butterfly (none for the unit twiddle), split by the GLV
// endomorphism into two interleaved 128-bit halves on G1.
#pragma once
#include "kernels.cuh"

namespace msm {

template <class F> MSM_COLD Xyzz<F> xyzz_neg(const Xyzz<F>& a) {
  Xyzz<F> r = a;
  r.y = F::norm(F::template sub<5, 1>(F::zero(), a.y));  // stored Y is below 4p in every field class
  return r;
}

// k * p for a canonical 256-bit k: signed 4-bit windows over the multiples 1p .. 8p
// (256 doublings + at most 65 additions; POINT_mul of ag-build/cl/ec.cl:122-130 is 256 + 256).
template <class F> MSM_COLD Xyzz<F> xyzz_scalar_mul(const Xyzz<F>& p, const uint32_t k[8]) {
  if (xyzz_is_inf<F>(p)) return p;
  Xyzz<F> T[8];
  T[0] = p;
  T[1] = xyzz_dbl<F>(p);
  for (int i = 2; i < 8; i++) T[i] = xyzz_add<F>(T[i - 1], p);
  int8_t d[65];
  uint32_t carry = 0;
  for (int i = 0; i < 64; i++) {
    const uint32_t raw = ((k[i >> 3] >> (4 * (i & 7))) & 15u) + carry;
    carry = raw > 8 ? 1u : 0u;
    d[i] = (int8_t)(carry ? (int)raw - 16 : (int)raw);
  }
  d[64] = (int8_t)carry;
  Xyzz<F> acc = xyzz_inf<F>();
  for (int i = 64; i >= 0; i--) {
    if (!xyzz_is_inf<F>(acc))
      for (int q = 0; q < 4; q++) acc = xyzz_dbl<F>(acc);
    const int di = d[i];
    if (di > 0) acc = xyzz_add<F>(acc, T[di - 1]);
    else if (di < 0) acc = xyzz_add<F>(acc, xyzz_neg<F>(T[-di - 1]));
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// GLV for the twiddle multiplications (G1 only).  Both curves have j = 0, so phi(x, y) = (beta x, y)
// with beta^3 = 1 in Fq acts on the r-torsion as multiplication by lambda (lambda^2 + lambda + 1 = 0
// mod r).  A twiddle k is split once, when the table is built, into k = k1 + k2 lambda with
// |k1|, |k2| < 2^128 (Babai rounding against the basis (a1, b1), (a2, b2) of the lattice
// {(x, y): x + y lambda = 0 mod r}: c_i = floor(k g_i / 2^320) with g1 = round(2^320 b2 / r),
// g2 = round(-2^320 b1 / r); k1 = k - c1 a1 - c2 a2, k2 = -c1 b1 - c2 b2).  The multiplication is then
// one interleaved double-and-add over 128 bits: 128 doublings + at most 66 additions instead of
// 256 + 65.  Constants derived with Python integers; checked against the oracle by
// every EC-FFT parity test (a wrong constant gives a wrong transform).
// ---------------------------------------------------------------------------------------------
struct GlvParams {
  uint64_t g1[4], g2[4];                  // multiply-shift constants, shift 320
  uint64_t a1[2], b1[2], a2[2], b2[2];    // magnitudes of the lattice basis
  uint32_t neg_a1, neg_b1, neg_a2, neg_b2;
  uint32_t beta[12];                      // beta in the API layout (Montgomery), N words used
};
inline GlvParams glv_params(bool bn254) {
  GlvParams p;
  memset(&p, 0, sizeof(p));
  if (bn254) {
    const uint64_t g1[4] = {0x6eb9c714773a6ef3ull, 0xd91d232ec7e0b3d7ull, 0x0000000000000002ull, 0x0ull};
    const uint64_t g2[4] = {0xa5e38cfb5eaa26daull, 0x7a7bd9d4391eb18dull, 0x4ccef014a773d2cfull, 0x0000000000000002ull};
    const uint64_t a1[2] = {0x89d3256894d213e3ull, 0x0ull}, b1[2] = {0x8211bbeb7d4f1128ull, 0x6f4d8248eeb859fcull};
    const uint64_t a2[2] = {0x0be4e1541221250bull, 0x6f4d8248eeb859fdull}, b2[2] = {0x89d3256894d213e3ull, 0x0ull};
    const uint32_t beta[8] = {0xd782e155u, 0x71930c11u, 0xffbe3323u, 0xa6bb947cu, 0xd4741444u, 0xaa303344u, 0x26594943u, 0x2c3b3f0du};
    memcpy(p.g1, g1, 32); memcpy(p.g2, g2, 32); memcpy(p.a1, a1, 16); memcpy(p.b1, b1, 16);
    memcpy(p.a2, a2, 16); memcpy(p.b2, b2, 16); memcpy(p.beta, beta, 32);
    p.neg_b1 = 1;
  } else {
    const uint64_t g1[4] = {0x389f49a7268bf7a4ull, 0x63f6e522f6cfee30ull, 0x7c6becf1e01faaddull, 0x0000000000000001ull};
    const uint64_t g2[4] = {0x355094edfede377cull, 0x0000000000000002ull, 0x0ull, 0x0ull};
    const uint64_t a1[2] = {0x00000000ffffffffull, 0xac45a4010001a402ull}, b1[2] = {0x1ull, 0x0ull};
    const uint64_t a2[2] = {0x1ull, 0x0ull}, b2[2] = {0x0000000100000000ull, 0xac45a4010001a402ull};
    const uint32_t beta[12] = {0x8671f071u, 0xcd03c9e4u, 0x1fcda5d2u, 0x5dab2246u, 0xd3851b95u, 0x587042afu,
                               0x01bacb9eu, 0x8eb60ebeu, 0x83d050d2u, 0x03f97d6eu, 0x54638741u, 0x18f02065u};
    memcpy(p.g1, g1, 32); memcpy(p.g2, g2, 32); memcpy(p.a1, a1, 16); memcpy(p.b1, b1, 16);
    memcpy(p.a2, a2, 16); memcpy(p.b2, b2, 16); memcpy(p.beta, beta, 48);
    p.neg_b1 = 1;
  }
  return p;
}

// floor(k g / 2^320) for 256-bit k, g: limbs 5 and 6 of the product (the result is below 2^128)
MSM_D void glv_mulshift(const uint64_t k[4], const uint64_t g[4], uint64_t c[2]) {
  uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; j++) {
      const uint64_t lo = k[i] * g[j], hi = __umul64hi(k[i], g[j]);
      uint64_t s = t[i + j] + lo;
      uint64_t c0 = s < lo;
      s += carry;
      c0 += s < carry;
      t[i + j] = s;
      carry = hi + c0;
    }
    t[i + 4] = carry;
  }
  c[0] = t[5];
  c[1] = t[6];
}
// acc (5 limbs, two's complement) += sign * c * m, c and m 128-bit magnitudes
MSM_D void glv_addmul(uint64_t acc[5], const uint64_t c[2], const uint64_t m[2], bool negative) {
  uint64_t t[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 2; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 2; j++) {
      const uint64_t lo = c[i] * m[j], hi = __umul64hi(c[i], m[j]);
      uint64_t s = t[i + j] + lo;
      uint64_t c0 = s < lo;
      s += carry;
      c0 += s < carry;
      t[i + j] = s;
      carry = hi + c0;
    }
    t[i + 2] += carry;
  }
  if (negative) {  // two's complement of t
    uint64_t carry = 1;
    for (int i = 0; i < 5; i++) {
      const uint64_t v = ~t[i] + carry;
      carry = (carry && v == 0) ? 1 : 0;
      t[i] = v;
    }
  }
  uint64_t carry = 0;
  for (int i = 0; i < 5; i++) {
    const uint64_t s = acc[i] + t[i];
    const uint64_t c0 = s < t[i];
    const uint64_t s2 = s + carry;
    carry = c0 + (s2 < carry);
    acc[i] = s2;
  }
}
// |v| of a 5-limb two's complement value known to fit 128 bits; returns the sign
MSM_D bool glv_abs(uint64_t v[5], uint32_t out[4]) {
  const bool negative = (v[4] >> 63) != 0;
  if (negative) {
    uint64_t carry = 1;
    for (int i = 0; i < 5; i++) {
      const uint64_t x = ~v[i] + carry;
      carry = (carry && x == 0) ? 1 : 0;
      v[i] = x;
    }
  }
  out[0] = (uint32_t)v[0]; out[1] = (uint32_t)(v[0] >> 32);
  out[2] = (uint32_t)v[1]; out[3] = (uint32_t)(v[1] >> 32);
  return negative;
}
// k (canonical, 8 words) -> out[0..3] = |k1|, out[4..7] = |k2|; returns bit 0: k1 < 0, bit 1: k2 < 0
MSM_D uint32_t glv_decompose(const GlvParams& gp, const uint32_t k[8], uint32_t out[8]) {
  uint64_t kk[4], c1[2], c2[2];
  for (int i = 0; i < 4; i++) kk[i] = ((uint64_t)k[2 * i + 1] << 32) | k[2 * i];
  glv_mulshift(kk, gp.g1, c1);
  glv_mulshift(kk, gp.g2, c2);
  uint64_t k1[5] = {kk[0], kk[1], kk[2], kk[3], 0}, k2[5] = {0, 0, 0, 0, 0};
  glv_addmul(k1, c1, gp.a1, !gp.neg_a1);  // k1 = k - c1 a1 - c2 a2
  glv_addmul(k1, c2, gp.a2, !gp.neg_a2);
  glv_addmul(k2, c1, gp.b1, !gp.neg_b1);  // k2 = -c1 b1 - c2 b2
  glv_addmul(k2, c2, gp.b2, !gp.neg_b2);
  uint32_t sign = glv_abs(k1, out) ? 1u : 0u;
  sign |= glv_abs(k2, out + 4) ? 2u : 0u;
  return sign;
}

// (+-k1) p + (+-k2) phi(p): one interleaved double-and-add over signed 4-bit windows of the two
// 128-bit halves; phi(T) = (beta X, Y, ZZ, ZZZ) costs one product when a table entry is used.
template <class F>
MSM_COLD Xyzz<F> xyzz_scalar_mul_glv(const Xyzz<F>& p, const uint32_t k[8], uint32_t sign, const typename F::Elem& beta) {
  if (xyzz_is_inf<F>(p)) return p;
  Xyzz<F> T[8];
  T[0] = p;
  T[1] = xyzz_dbl<F>(p);
  for (int i = 2; i < 8; i++) T[i] = xyzz_add<F>(T[i - 1], p);
  int8_t d[2][33];
  for (int h = 0; h < 2; h++) {
    uint32_t carry = 0;
    for (int i = 0; i < 32; i++) {
      const uint32_t raw = ((k[4 * h + (i >> 3)] >> (4 * (i & 7))) & 15u) + carry;
      carry = raw > 8 ? 1u : 0u;
      d[h][i] = (int8_t)(carry ? (int)raw - 16 : (int)raw);
    }
    d[h][32] = (int8_t)carry;
  }
  Xyzz<F> acc = xyzz_inf<F>();
  for (int i = 32; i >= 0; i--) {
    if (!xyzz_is_inf<F>(acc))
      for (int q = 0; q < 4; q++) acc = xyzz_dbl<F>(acc);
    for (int h = 0; h < 2; h++) {
      int di = d[h][i];
      if (di == 0) continue;
      if ((sign >> h) & 1) di = -di;
      Xyzz<F> t = T[(di > 0 ? di : -di) - 1];
      if (h) t.x = F::norm(F::mul(t.x, beta));
      acc = xyzz_add<F>(acc, di > 0 ? t : xyzz_neg<F>(t));
    }
  }
  return acc;
}

// tw[j] = omega^j as a canonical little-endian integer, j < half_n (tw_sign != nullptr: split for GLV).
// omegas[i] = omega^(2^i) in Montgomery form (arkworks' in-memory Fr; ag-cuda-ec/src/ec_fft.rs:120-124
// builds the same array).
template <class PR>
__global__ void k_fft_twiddles(const uint32_t* __restrict__ omegas, uint32_t half_n, uint32_t* __restrict__ tw,
                               GlvParams gp, uint8_t* __restrict__ tw_sign) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= half_n) return;
  Fp<PR> acc = fp_one<PR>();
  for (uint32_t i = 0; (j >> i) != 0; i++) {
    if ((j >> i) & 1) {
      Fp<PR> w;
      uint32_t k[8];
      load_scalar(omegas, i, k);
#pragma unroll
      for (int q = 0; q < 8; q++) w.v[q] = k[q];
      acc = fp_mul<PR>(acc, w);
    }
  }
  const Fp<PR> c = fp_from_mont<PR>(acc);
  uint32_t o8[8];
#pragma unroll
  for (int q = 0; q < 8; q++) o8[q] = c.v[q];
  if (tw_sign) {  // G1: store |k1| | |k2| and the two signs (GLV, see below)
    uint32_t dec[8];
    tw_sign[j] = (uint8_t)glv_decompose(gp, o8, dec);
#pragma unroll
    for (int q = 0; q < 8; q++) o8[q] = dec[q];
  }
  uint4* o = reinterpret_cast<uint4*>(tw) + 2 * (size_t)j;
  o[0] = make_uint4(o8[0], o8[1], o8[2], o8[3]);
  o[1] = make_uint4(o8[4], o8[5], o8[6], o8[7]);
}

MSM_HD uint32_t bit_reverse(uint32_t v, uint32_t bits) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < bits; i++) {
    r = (r << 1) | (v & 1);
    v >>= 1;
  }
  return r;
}

// API Jacobian -> XYZZ work array in bit-reversed order (the permutation serial_ec_fft starts with,
// ec-gpu-proxy/src/ec_fft_cpu.rs:27-32)
template <class F>
__global__ void k_fft_load(const ApiJacobian<F>* __restrict__ in, uint32_t log_n, Xyzz<F>* __restrict__ x) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << log_n)) return;
  store_vec(&x[bit_reverse(i, log_n)], xyzz_from_api_jacobian<F>(&in[i]));
}
template <class F>
__global__ void k_fft_store(const Xyzz<F>* __restrict__ x, uint32_t n, ApiJacobian<F>* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xyzz_to_api_jacobian<F>(load_vec(&x[i]), &out[i]);
}

// One decimation-in-time round with half-size m: butterfly (k + j, k + j + m), twiddle
// w_m^j = omega^(j n / 2m) = tw[j * tw_stride]   (ec-gpu-proxy/src/ec_fft_cpu.rs:35-54)
template <class F>
__global__ void __launch_bounds__(64)
k_fft_round(Xyzz<F>* __restrict__ x, uint32_t n, uint32_t m, uint32_t tw_stride, const uint32_t* __restrict__ tw,
            const uint8_t* __restrict__ tw_sign, ApiElem<F> beta_api) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n / 2) return;
  const uint32_t j = b & (m - 1);
  const uint32_t i0 = ((b - j) << 1) + j, i1 = i0 + m;
  const Xyzz<F> lo = load_vec(&x[i0]);
  Xyzz<F> hi = load_vec(&x[i1]);
  if (j) {
    uint32_t k[8];
    load_scalar(tw, j * tw_stride, k);
    if (tw_sign) hi = xyzz_scalar_mul_glv<F>(hi, k, tw_sign[j * tw_stride], F::from_api(beta_api.w));
    else hi = xyzz_scalar_mul<F>(hi, k);
  }
  store_vec(&x[i0], xyzz_add<F>(lo, hi));
  store_vec(&x[i1], xyzz_add<F>(lo, xyzz_neg<F>(hi)));
}

}  // namespace msm
