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
// multiplication in XYZZ coordinates per butterfly (none for the unit twiddle).
#pragma once
#include "kernels.cuh"

namespace msm {

// tw[j] = omega^j as a canonical little-endian integer, j < half_n.  omegas[i] = omega^(2^i) in
// Montgomery form (arkworks' in-memory Fr; ag-cuda-ec/src/ec_fft.rs:120-124 builds the same array).
template <class PR>
__global__ void k_fft_twiddles(const uint32_t* __restrict__ omegas, uint32_t half_n, uint32_t* __restrict__ tw) {
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
  uint4* o = reinterpret_cast<uint4*>(tw) + 2 * (size_t)j;
  o[0] = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
  o[1] = make_uint4(c.v[4], c.v[5], c.v[6], c.v[7]);
}

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
k_fft_round(Xyzz<F>* __restrict__ x, uint32_t n, uint32_t m, uint32_t tw_stride, const uint32_t* __restrict__ tw) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n / 2) return;
  const uint32_t j = b & (m - 1);
  const uint32_t i0 = ((b - j) << 1) + j, i1 = i0 + m;
  const Xyzz<F> lo = load_vec(&x[i0]);
  Xyzz<F> hi = load_vec(&x[i1]);
  if (j) {
    uint32_t k[8];
    load_scalar(tw, j * tw_stride, k);
    hi = xyzz_scalar_mul<F>(hi, k);
  }
  store_vec(&x[i0], xyzz_add<F>(lo, hi));
  store_vec(&x[i1], xyzz_add<F>(lo, xyzz_neg<F>(hi)));
}

}  // namespace msm
